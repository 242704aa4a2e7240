"""Writes profiles/r2_sass_tma_excerpt.txt: the TMA / mbarrier / warp-shuffle instructions in the SASS of the
shipped library, per kernel (cuobjdump -sass; works without a GPU)."""
import hashlib
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "zetaray_b200", "libzetaray_b200.so")
PAT = re.compile(r"UTMALDG|UTMASTG|UBLKCP|SYNCS|UTMACMDFLUSH|UTMAPF|SHFL|FENCE.VIEW.ASYNC")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    fn, cnt, ex = None, {}, {}
    for l in sass.split("\n"):
        m = re.search(r"Function : (\S+)", l)
        if m:
            fn = m.group(1)
            continue
        if fn and PAT.search(l):
            key = re.sub(r"^\s*/\*[0-9a-f]+\*/\s*", "", l).split(";")[0].strip()
            op = key.split()[1] if key.startswith("@") else key.split()[0]
            cnt.setdefault(fn, {}).setdefault(op, 0)
            cnt[fn][op] += 1
            ex.setdefault(fn, {}).setdefault(op, key)
    sha = hashlib.sha256(open(SO, "rb").read()).hexdigest()
    out = os.path.join(ROOT, "profiles", "r2_sass_tma_excerpt.txt")
    with open(out, "w") as f:
        f.write("# cuobjdump -sass zetaray_b200/libzetaray_b200.so: TMA / mbarrier / warp-shuffle instructions per kernel\n")
        f.write("# (regenerate: python tools/sass_excerpt.py). library sha256 %s\n" % sha)
        for fn in sorted(cnt):
            c = cnt[fn]
            if not any(k.startswith(("UTMA", "UBLKCP", "SYNCS")) for k in c):
                continue
            d = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
            f.write("\n%s\n" % d[:200])
            for k in sorted(c):
                f.write("   %4d x %-30s e.g. %s\n" % (c[k], k, ex[fn][k]))
    print("wrote", out)


if __name__ == "__main__":
    main()
