"""Writes profiles/sass/<kernel>.sass for every kernel of the built library: `cuobjdump -sass` of
jaxsplat_b200/lib/libjaxsplat_b200.so with the encodings stripped.  For templated kernels the
instantiation the default configuration launches is kept (PREFER below).
usage: python tools/dump_sass.py [lib.so] [outdir]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "jaxsplat_b200", "lib", "libjaxsplat_b200.so")
OUT = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "sass")
# kernel base name -> substring of the mangled template arguments of the default instantiation
PREFER = {"raster_fwd16x2_kernel": "ILb1E", "raster_bwd16x2p_kernel": "ILb1ELb0E",
          "raster_bwd_hits_kernel": "ILb0ELb1ELb1ELb0ELb0E", "project_fwd_kernel": "ILb1E",
          "project_bwd_kernel": "ILb1ELb1E", "raster_fwd_warp_kernel": "ILb1ELb1ELb0E",
          "rc_col_scatter_kernel": "ILb1E", "mse_loss_grad_kernel": "IhE",
          "onesweep_pass_kernel": "IjLi16ELb1E", "radix_hist_kernel": "IjE", "upsweep_kernel": "IjE"}


def main():
    text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    os.makedirs(OUT, exist_ok=True)
    blocks = re.split(r"\n\s*Function : ", text)[1:]
    chosen = {}
    for blk in blocks:
        mangled = blk.split("\n", 1)[0].strip()
        base = None
        for m in re.finditer(r"[a-z][a-z0-9_]*?_kernel", mangled):
            for i in range(m.start(), m.end() - 7):  # Itanium mangling: <length><name>
                name = mangled[i:m.end()]
                if name[0].isalpha() and mangled[:i].endswith(str(len(name))):
                    base = name
                    break
            if base:
                break
        if not base:
            continue
        want = PREFER.get(base)
        if base in chosen and not (want and want in mangled.split(base, 1)[1][:24]):
            continue
        lines = []
        for ln in blk.split("\n")[1:]:
            mm = re.match(r"\s+(/\*[0-9a-f]{4}\*/)\s+(.*?);\s*/\* 0x[0-9a-f]+ \*/", ln)
            if mm:
                lines.append(f"{mm.group(1)} {mm.group(2).rstrip()}  ;")
        chosen[base] = (mangled, lines)
    for base, (mangled, lines) in sorted(chosen.items()):
        with open(os.path.join(OUT, base + ".sass"), "w") as f:
            f.write(f"// {mangled}\n// sm_100a SASS (cuobjdump -sass, encodings stripped)\n")
            f.write("\n".join(lines) + "\n")
        print(f"{base:34s} {len(lines):5d} instructions")


if __name__ == "__main__":
    main()
