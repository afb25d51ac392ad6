#!/usr/bin/env python
"""Is the built extension the code the evidence shows?  Compares the instruction stream of every kernel in
atomo_b200/_C*.so (cuobjdump -sass) with the committed listing under profiles/sass/ (taken from the build that was
validated and profiled on the B200).  Exit code 1 on any difference; run after touching a .cu file, refresh the listings
with scripts/gpu_evidence.sh when the change is intended."""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    so = glob.glob(os.path.join(ROOT, "atomo_b200", "_C*.so"))
    if not so:
        print("no built extension (python setup.py build_ext --inplace)")
        return 2
    sass = subprocess.run(["cuobjdump", "-sass", so[0]], capture_output=True, text=True, check=True).stdout
    funcs = {}
    for part in re.split(r"\n\s*Function : ", sass)[1:]:
        head, _, body = part.partition("\n")
        ins = [re.sub(r"\s+", " ", m.group(1)).strip()
               for m in re.finditer(r"^\s*/\*[0-9a-f]{4,5}\*/\s+(.*?)\s*;", body, re.M)]
        funcs[head.strip()] = ins
    bad = 0
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "sass", "*.sass.txt"))):
        name = os.path.basename(f)[:-len(".sass.txt")]
        v2 = name.startswith("v2_v2_")
        kernel = name[3:] if v2 else name
        ref = [re.sub(r"\s+", " ", l).strip().rstrip(";").strip() for l in open(f) if l.strip() and not l.startswith("//")]
        cands = [k for k in funcs if kernel in k and (("2v2" in k) == v2)]
        cur = funcs[cands[0]] if cands else None
        if cur is None:
            print("%-34s not in the binary" % name)
            bad += 1
        elif cur != ref:
            n = next((i for i in range(min(len(cur), len(ref))) if cur[i] != ref[i]), min(len(cur), len(ref)))
            print("%-34s DIFFERS at instruction %d (%d committed, %d built)" % (name, n, len(ref), len(cur)))
            bad += 1
        else:
            print("%-34s %5d instructions identical" % (name, len(ref)))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
