"""SASS evidence for profiles/: per conv / wgrad kernel the counts of the tensor-core, TMEM and TMA instructions
(UTCMMA / UTCHMMA, LDTM, UTMALDG / UBLKCP, UTCBAR, SYNCS) and, for the hot kernels of the C2 step, the MMA issue block
verbatim.  usage: cuobjdump -sass simplesr_b200/libssr_b200.so > /tmp/sass.txt; python tools/sass_listing.py /tmp/sass.txt out.txt"""
import collections
import re
import sys

MNEMONICS = ["UTCHMMA", "UTCMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS",
             "ELECT", "ACQBULK", "UTMAPF", "R2UR", "NANOSLEEP", "STG", "LDG", "STS", "LDS", "HMMA"]
HOT = ["conv_tc_kernelILi3ELi72ELb1E", "conv_tc_kernelILi3ELi17ELb1E", "conv_tc_kernelILi3ELi33ELb0E",
       "conv_tc_kernelILi3ELi17ELb0E", "wgrad_tc"]

funcs = collections.OrderedDict()
cur = None
for line in open(sys.argv[1], errors="replace"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    if cur is not None and re.search(r"/\*[0-9a-f]{4}\*/", line):
        funcs[cur].append(line.rstrip())

out = open(sys.argv[2], "w")
out.write("# cuobjdump -sass simplesr_b200/libssr_b200.so (sm_100a), reduced by tools/sass_listing.py\n")
out.write("# kernel, instructions, " + ", ".join(MNEMONICS) + "\n")
for name, body in funcs.items():
    if not any(k in name for k in ("conv_tc", "wgrad", "comm_", "adam", "dense_")):
        continue
    cnt = collections.Counter()
    for l in body:
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            op = m.group(1)
            for k in MNEMONICS:
                if op.startswith(k):
                    cnt[k] += 1
                    break
    out.write(f"{name}, {len(body)}, " + ", ".join(str(cnt[k]) for k in MNEMONICS) + "\n")
out.write("\n")
for key in HOT:
    for name, body in funcs.items():
        if key not in name:
            continue
        idx = [i for i, l in enumerate(body) if "UTCHMMA" in l or "UTCMMA" in l]
        if not idx:
            continue
        # the longest run of closely spaced MMAs = one fully unrolled (tap, k-step) block
        best, start = (0, 0), idx[0]
        prev = idx[0]
        for i in idx[1:] + [10 ** 9]:
            if i - prev > 8:
                if prev - start > best[1] - best[0]:
                    best = (start, prev)
                start = i
            prev = i
        lo, hi = best
        hi = min(hi, lo + 60)
        out.write(f"## {name}: MMA issue block, SASS lines {lo}..{hi} of {len(body)}\n")
        for l in body[max(0, lo - 6):hi + 4]:
            out.write(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l) + "\n")
        out.write("\n")
        break
out.close()
