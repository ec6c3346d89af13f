"""Builds tf_shim/libssr_tf_ops.so against the installed TensorFlow and the in-tree libssr_b200.so.
Skips LOUDLY (exit code 0, message on stderr) when TensorFlow cannot be imported - the case in this repository's image."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def main():
    try:
        import tensorflow as tf
    except Exception as e:   # noqa: BLE001
        print(f"tf_shim: SKIPPED - TensorFlow is not importable here ({type(e).__name__}: {e}); the shim source "
              "(ssr_tf_ops.cc, ssr_tf.py) is shipped untested", file=sys.stderr)
        return 0
    lib_dir = os.path.join(ROOT, "simplesr_b200")
    cmd = (["g++", "-std=c++14", "-shared", "-fPIC", "-O2", os.path.join(HERE, "ssr_tf_ops.cc"), "-o",
            os.path.join(HERE, "libssr_tf_ops.so"), "-I/usr/local/cuda/include", "-DGOOGLE_CUDA=1"]
           + tf.sysconfig.get_compile_flags() + tf.sysconfig.get_link_flags()
           + [f"-L{lib_dir}", "-lssr_b200", f"-Wl,-rpath,{lib_dir}", "-L/usr/local/cuda/lib64", "-lcudart"])
    print(" ".join(cmd))
    return subprocess.call(cmd)


if __name__ == "__main__":
    sys.exit(main())
