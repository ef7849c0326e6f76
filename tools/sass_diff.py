#!/usr/bin/env python
"""sass_diff.py — are the kernels of two builds the same machine code?

    cuobjdump -sass old/libgcz_b200.so > old.sass ; cuobjdump -sass gecoz_b200/libgcz_b200.so > new.sass
    python tools/sass_diff.py old.sass new.sass [--map 'REGEX=>REPLACEMENT' ...] [-v]

Compares the SASS of every kernel of the OLD dump with the kernel of the same demangled name in the NEW dump (function by
function, instruction text and encodings).  --map rewrites NEW names before matching (a kernel that gained a defaulted template
parameter keeps its code but not its name: --map '(onesweep_kernel<[^>]*), 8>=>\\1>').  Used to show that experimental variants added
behind environment variables leave the kernels that were measured untouched.  Needs cuobjdump and c++filt, no GPU.
"""
import argparse
import difflib
import hashlib
import re
import subprocess


def kernels(path):
    text = open(path).read()
    out = {}
    for part in re.split(r"\n\s*Function : ", text)[1:]:
        name, _, body = part.partition("\n")
        body = body.split("\nFatbin elf code")[0].rstrip()
        dem = subprocess.run(["c++filt", name.strip()], capture_output=True, text=True).stdout.strip()
        out[dem] = body
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("old")
    ap.add_argument("new")
    ap.add_argument("--map", action="append", default=[])
    ap.add_argument("-v", action="store_true")
    args = ap.parse_args()
    old, new = kernels(args.old), kernels(args.new)
    renamed = {}
    for name, body in new.items():
        for m in args.map:
            pat, _, rep = m.partition("=>")
            name = re.sub(pat, rep, name)
        renamed[name] = body
    same = differ = missing = 0
    for name, body in old.items():
        if name not in renamed:
            missing += 1
            print("MISSING", name[:160])
        elif hashlib.md5(body.encode()).digest() == hashlib.md5(renamed[name].encode()).digest():
            same += 1
        else:
            differ += 1
            print("DIFFERS", name[:160])
            if args.v:
                d = list(difflib.unified_diff(body.splitlines(), renamed[name].splitlines(), lineterm="", n=0))
                print("\n".join(d[:40]))
    print(f"{len(old)} kernels in the old build: {same} identical, {differ} differ, {missing} not found by name; {len(new) - len(old)} more in the new build")


if __name__ == "__main__":
    main()
