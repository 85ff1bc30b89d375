"""profiles/<round>_sass_mnemonics.txt: which tensor / TMA / barrier SASS instructions each kernel of libcdm_b200.so
contains (cuobjdump -sass, runs without a GPU).

    python tools/sass_mnemonics.py r1
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "camels-diffusion-model_b200", "libcdm_b200.so")
KEEP = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "HMMA", "SYNCS", "REDUX", "RED", "ATOM", "ATOMG")
HEAD = """SASS mnemonics per kernel of libcdm_b200.so (cuobjdump -sass, sm_100a), tensor/TMA-relevant ones only:
UTCHMMA = tcgen05.mma (kind::f16), UTMALDG = TMA tensor load, UTMASTG = TMA tensor store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
HMMA = mma.sync (first/last convolution), SYNCS = mbarrier ops, RED/ATOM = global reductions / atomics.
"""


def main(tag):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = {}
    for m in re.finditer(r"Function : (\S+)", sass):
        names[m.group(1)] = None
    dem = subprocess.run(["cu++filt"] + list(names), capture_output=True, text=True, check=True).stdout.splitlines()
    pretty = dict(zip(names, dem))
    out, cur, cnt = [HEAD], None, None
    per = collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\((?:int|bool)\)", "", pretty[m.group(1)])  # conv3x3_kernel<(int)2>(...) -> conv3x3_kernel<2>
            cur = re.sub(r"\(.*", "", cur)
            cnt = per.setdefault(cur, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9]+)", line)
        if m and cur and m.group(1) in KEEP:
            cnt[m.group(1)] += 1
    for k, c in per.items():
        if c:
            out.append(f"{k}: " + ", ".join(f"{n} x{v}" for n, v in sorted(c.items())))
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_mnemonics.txt")
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")
    print("wrote", path, len(per), "kernels")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r1")
