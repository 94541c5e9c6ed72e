"""cuobjdump -sass of the shipped library -> per-kernel opcode histogram (profiles/r02_sass_summary.txt).
Shows what the kernels are made of: 128-bit global loads/stores, shared-memory table lookups, warp shuffles / redux,
atomics of the tile scheduler; no tensor-core instruction (UTC*MMA / HMMA): nothing here is a contraction.  The bulk-copy engine
(UBLKCP = cp.async.bulk) appears in exactly one kernel, xxh64_tma_kernel, an opt-in A/B that measured slower than the plain kernel;
for the pull over PCIe the same engine did not beat plain loads either (tools/h2d_probe.cu, modes 5-7)."""
import collections, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else "light-speed-io_b200/liblsio_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
demangle = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
print(f"# {lib}: {len(kernels)} sm_100a kernels (cuobjdump -sass); instruction counts are static (code size), not dynamic")
groups = [("global 128-bit loads", r"^LDG\.E\..*128"), ("global 128-bit stores", r"^STG\.E\..*128"), ("other global loads", r"^LDG"), ("other global stores", r"^STG"),
          ("shared loads", r"^LDS"), ("shared stores", r"^STS"), ("constant/param loads", r"^LDC|^ULDC"), ("atomics / reductions", r"^ATOM|^RED|^REDUX|^ATOMG|^ATOMS"),
          ("warp shuffles / votes", r"^SHFL|^VOTE|^MATCH"), ("logic (LOP3 / SHF / PRMT)", r"^LOP3|^SHF|^PRMT|^ULOP3|^USHF|^UPRMT"), ("integer add / mad", r"^IADD|^IMAD|^LEA|^UIADD|^UIMAD|^ULEA|^VIADD"),
          ("barriers / fences", r"^BAR|^MEMBAR|^FENCE|^SYNCS|^ERRBAR"), ("TMA / bulk copy", r"^UTMA|^UBLKCP"), ("tensor core", r"MMA|^UTC")]
for (name, cnt), pretty in zip(kernels.items(), demangle):
    total = sum(cnt.values())
    print(f"\n## {pretty}\n   {total} instructions")
    used = collections.Counter()
    for g, pat in groups:
        n = 0
        for op, c in cnt.items():
            if op not in used and re.search(pat, op):
                n += c
                used[op] = 1
        print(f"   {g:28s} {n}")
    top = ", ".join(f"{op} {c}" for op, c in cnt.most_common(8))
    print(f"   top opcodes: {top}")
