"""SASS evidence of what the built library really contains: per kernel family, the count of the instructions that
prove the sm_100a paths (tcgen05: UTCHMMA / LDTM / UTCBAR, TMA bulk copies: UBLKCP, mbarrier: SYNCS, packed fp32:
FADD2 / FMUL2, system-scope release / acquire of the halo flags).   python tools/sass_histogram.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fluidnet_cxx_b200", "_lib", "libfluidstep_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fam_of = [("k_conv_tc", "k_conv_tc (tcgen05 implicit-GEMM convolutions, conv_tc.cu)"),
          ("k_jacobi2d_blocked", "k_jacobi2d_blocked (temporally blocked Jacobi, jacobi_blocked.cu)"),
          ("k_halo_exchange", "k_halo_exchange (peer-memory halo exchange, halo.cu)"),
          ("k2_advect_clean", "k2_advect_clean (fused MacCormack advection, interior fast path, step2d.cu)"),
          ("k2_advect", "k2_advect (fused MacCormack advection, generic path, step2d.cu)"),
          ("k2_forces_div", "k2_forces_div (step2d.cu)"), ("k2_project", "k2_project / k2_project4 (step2d.cu)"),
          ("k_output_fields", "k_output_fields (drivers' output block, stencils.cu)")]
keys = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "FADD2", "FMUL2", "FFMA2", "FADD", "FMUL", "FFMA", "SHFL", "LDS", "STS",
        "LDG", "STG", "BAR", "MEMBAR", "ATOM", "RED"]
counts = collections.defaultdict(collections.Counter)
nk = collections.Counter()
strong = collections.defaultdict(collections.Counter)
fam = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fam = next((label for key, label in fam_of if key in m.group(1)), None)
        if fam:
            nk[fam] += 1
        continue
    if fam is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if not m:
        continue
    op, mods = m.group(1), m.group(2)
    if op in keys:
        counts[fam][op] += 1
    if ".SYS" in mods and op in ("ST", "LD", "STG", "LDG", "MEMBAR", "FENCE", "ATOM", "ATOMG", "RED"):
        strong[fam][op + mods] += 1
print(f"# {os.path.relpath(so, ROOT)}: SASS opcode counts per kernel family (cuobjdump -sass, all template instantiations summed)")
for key, label in fam_of:
    if label not in nk:
        continue
    c = counts[label]
    print(f"\n{label}: {nk[label]} kernels")
    print("  " + ", ".join(f"{k} {c[k]}" for k in keys if c[k]))
    if strong[label]:
        print("  system-scope memory operations: " + ", ".join(f"{k} x{v}" for k, v in strong[label].most_common()))
