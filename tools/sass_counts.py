"""python tools/sass_counts.py > profiles/r02_sass_counts.md : SASS mnemonic counts of the hot kernels (cuobjdump -sass on the in-tree
objects): which hardware paths each kernel uses (tcgen05 = UTCHMMA / LDTM / UTCBAR, TMA = UTMALDG / UBLKCP, packed fp32 = FFMA2 / FMUL2,
L2 reductions = REDG, DSMEM = ST.ASYNC / mapa-based stores, mbarrier = SYNCS)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = os.path.join(ROOT, "rsrl_b200", "csrc", "build")
TARGETS = [("inst_f32_d0.o", r"persistent_kernelIfLi0ELi0ELi5ELi3ELi0E", "persistent_kernel<float, MountainCar, Fourier, 5, 3, SHARED> (cfg2 headline)"),
           ("inst_f32_d0.o", r"persistent_kernelIfLi0ELi0ELi5ELi3ELi2E", "persistent_kernel<float, ..., SHARED + traces> (cfg5)"),
           ("inst_f32_d0.o", r"persistent_kernelIfLi0ELi0ELi5ELi3ELi1E", "persistent_kernel<float, ..., PER_ENV>"),
           ("inst_f64_d0.o", r"persistent_kernelIdLi0ELi0ELi5ELi3ELi0E", "persistent_kernel<double, ..., SHARED> (cluster + DSMEM exchange)"),
           ("tile_f32.o", r"tile_dense_kernelIfLi1ELi2ELb0ELi1024ELi8E", "tile_dense_kernel<float, CartPole, 2, false, 1024, 8> (cfg3, launched with 896 threads)"),
           ("f4tc.o", r"f4tc_q_kernelILi2ELi0ELb0E", "f4tc_q_kernel<Acrobot, 0> (cfg4: Q(s))"),
           ("f4tc.o", r"f4tc_q_kernelILi2ELi1ELb0E", "f4tc_q_kernel<Acrobot, 1> (cfg4: Q(s') + TD)"),
           ("f4tc.o", r"f4tc_dw_kernelILi2E", "f4tc_dw_kernel<Acrobot> (cfg4: dW)"),
           ("f4tc.o", r"f4tc_phys_kernelILi2E", "f4tc_phys_kernel<Acrobot>"),
           ("f4tc.o", r"f4tc_reduce_kernel", "f4tc_reduce_kernel")]
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FFMA", "DFMA", "LDS", "STS", "REDG", "ATOMG",
        "SHFL", "BAR", "LDL", "STL"]
print("# SASS mnemonic counts (static) of the hot kernels — `python tools/sass_counts.py`, CUDA 12.9, sm_100a\n")
print("| kernel | registers | " + " | ".join(KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
for obj, pat, label in TARGETS:
    path = os.path.join(B, obj)
    names = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    fn = next((m.group(1) for m in re.finditer(r"Function : (\S+)", names) if re.search(pat, m.group(1))), None)
    if not fn:
        print(f"| {label} | not found | |"); continue
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, path], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    reg = re.search(re.escape(fn) + r":\s*\n\s*REG:(\d+)", res)
    cnt = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
        if m: cnt[m.group(1)] += 1
    print(f"| {label} | {reg.group(1) if reg else '?'} | " + " | ".join(str(cnt.get(k, 0)) for k in KEYS) + " |")
print("\nNo `UTMALDG` / `UBLKCP` anywhere: operand tiles are generated on chip by the CTA's threads (nothing to fetch from HBM but 32 B of state per env); "
      "the bulk-copy / bulk-reduction experiments of round 2 (`cp.async.bulk` DSMEM hops, `cp.reduce.async.bulk.add.u64` = `UBLKRED`) were measured "
      "slower and removed (profiles/r02_persistent.md).")
