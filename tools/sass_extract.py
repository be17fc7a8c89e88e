"""Write SASS listings of the hot kernels to profiles/sass/ (evidence that each variant uses the
instructions it claims: UTCHMMA / UTMALDG / LDTM for 3xTF32, FFMA + LDS.128 + UTMALDG for the
TMA-fed FFMA kernel, DMMA.8x8x4, DFMA).  Encodings are stripped to keep the files small.

    python tools/sass_extract.py
"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "openmp-blas_b200" / "libb200mtm.so"
OUT = ROOT / "profiles" / "sass"

KERNELS = {
    "tf32x3_2cta": r"mtm_tf32x3_kernelILi2ELb0ELb0E",
    "tf32x3_2cta_dyn": r"mtm_tf32x3_kernelILi2ELb1ELb0E",
    "tf32x3_1cta": r"mtm_tf32x3_kernelILi1ELb0ELb0E",
    "tf32x3_2cta_fused": r"mtm_tf32x3_kernelILi2ELb0ELb1E",
    "split_lo_planes": r"split_kernelILb0E",
    "ffma_tma_128x128x32_s3": r"mtm_ffma_tma_kernelILi128ELi32ELi3ELb0E",
    "ffma2_tma_128x128x32_s3": r"mtm_ffma_tma_kernelILi128ELi32ELi3ELb1E",
    "mtv_icontig_f32_v16": r"mtv_icontig_kernelIfLi4ELb1E",
    "mtv_kcontig_f32_v16_warp": r"mtv_kcontig_kernelIfLi4ELi1ELb0E",
    "transpose_vec_f32": r"transpose_vec_kernelIfLb1ELb1E",
    "dmma_64x64x8_w2x2_mode10": r"mtm_dmma_kernelILi64ELi64ELi8ELi2ELi2ELi4ELi1ELi0E",
    "dmma_tma_64x64x16_s3": r"mtm_dmma_tma_kernelILi3ELi4E",
    "dfma_128x128x8_t8x8_mode10": r"mtm_simt_kernelIdLi128ELi128ELi8ELi8ELi8ELi1ELi1ELi0E",
    "ffma_128x128x16_t8x8_mode10": r"mtm_simt_kernelIfLi128ELi128ELi16ELi8ELi8ELi2ELi1ELi0E",
}


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    index = []
    for name, pat in KERNELS.items():
        body = next((f for f in funcs if re.search(pat, f.split("\n", 1)[0])), None)
        if body is None:
            print("not found:", name, file=sys.stderr)
            continue
        lines = body.split("\n")
        mangled = lines[0].strip()
        text, ops = [], Counter()
        for l in lines[1:]:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", l)
            if m:
                ins = m.group(2).rstrip()
                text.append(f"/*{m.group(1)}*/ {ins} ;")
                tok = ins.split()
                op = tok[1] if tok[0].startswith("@") else tok[0]
                ops[op.split(".")[0] if not op.startswith(("UTC", "UTMA", "LDTM", "DMMA", "LDS", "LDG", "STS", "STG", "SYNCS")) else op] += 1
            elif "Fatbin" in l:
                break
        key = {k: v for k, v in ops.items() if k.startswith(("UTC", "UTMA", "LDTM", "DMMA", "DFMA", "FFMA", "LDS.128", "SYNCS"))}
        hist = "KEY " + ", ".join(f"{k} x{v}" for k, v in sorted(key.items())) + " | TOP " + \
               ", ".join(f"{k} x{v}" for k, v in ops.most_common(12))
        (OUT / f"{name}.sass").write_text(f"// {mangled}\n// arch sm_100a, {len(text)} instructions\n// mnemonics: {hist}\n" + "\n".join(text) + "\n")
        index.append(f"{name}: {len(text)} instr; {hist}")
    (OUT / "INDEX.txt").write_text("\n".join(index) + "\n")
    print("\n".join(index))


if __name__ == "__main__":
    main()
