"""Stand-alone GPU probe of the tcgen05 GEMM (run under gpurun; prints one line per configuration).

Not a pytest file: each group runs in its own process so that a fault in one configuration (which poisons the
CUDA context) cannot mask the others.  `python tests/gpu_probe_gemm.py --group cta1`
"""
import argparse
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from edward2_b200 import ops, _lib


def ref_gemm(a, b, major_a, major_b):
    A = a.double() if major_a == ops.MAJOR_K else a.double().t()
    B = b.double().t() if major_b == ops.MAJOR_K else b.double()
    return A @ B


def make_operands(m, n, k, major_a, major_b, dtype, gen):
    a = torch.randn((m, k) if major_a == ops.MAJOR_K else (k, m), generator=gen, device="cuda").to(dtype)
    b = torch.randn((n, k) if major_b == ops.MAJOR_K else (k, n), generator=gen, device="cuda").to(dtype)
    return a, b


def block_report(err, tag):
    """Where is the error? max |err| per 32-row x 64-col block, to recognise layout / swizzle bugs."""
    m, n = err.shape
    rows = []
    for r0 in range(0, min(m, 256), 32):
        cols = []
        for c0 in range(0, min(n, 512), 64):
            cols.append(f"{err[r0:r0+32, c0:c0+64].max().item():8.2e}")
        rows.append(" ".join(cols))
    print(f"    [{tag}] block max-abs-err map (32 rows x 64 cols):")
    for r in rows:
        print("      " + r)


def run_case(name, m, n, k, major_a, major_b, dtype=torch.bfloat16, tol=None, **kw):
    gen = torch.Generator(device="cuda").manual_seed(1234)
    a, b = make_operands(m, n, k, major_a, major_b, dtype, gen)
    ref = ref_gemm(a, b, major_a, major_b)
    try:
        out = ops.gemm(a, b, m, n, k, major_a=major_a, major_b=major_b, out_dtype=torch.float32, **kw)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"FAIL {name}: exception {e}")
        return False
    err = (out.double() - ref).abs()
    scale = ref.abs().max().item()
    rel = err.max().item() / max(scale, 1e-30)
    tol = tol if tol is not None else (2e-3 if dtype == torch.bfloat16 else 1e-5)
    ok = rel <= tol and bool(torch.isfinite(out).all())
    print(f"{'ok  ' if ok else 'FAIL'} {name}: m={m} n={n} k={k} rel_err={rel:.3e} (tol {tol:.0e})")
    if not ok:
        block_report(err, name)
    return ok


MAJ = {"K": ops.MAJOR_K, "MN": ops.MAJOR_MN}


def group_cta(cg):
    ok = True
    for ma in ("K", "MN"):
        for mb in ("K", "MN"):
            for bn in (128, 256):
                ok &= run_case(f"cta{cg} A={ma} B={mb} bn={bn} aligned", 256, 512, 256, MAJ[ma], MAJ[mb],
                               block_n=bn, cta_group=cg)
    # multi-tile, persistent loop, phases wrapping
    ok &= run_case(f"cta{cg} big", 1024, 1024, 1024, MAJ["K"], MAJ["MN"], block_n=256, cta_group=cg)
    ok &= run_case(f"cta{cg} big few-ctas", 1024, 1024, 512, MAJ["K"], MAJ["MN"], block_n=256, cta_group=cg, max_ctas=4)
    # ragged edges (TMA OOB fill / clipped stores)
    ok &= run_case(f"cta{cg} ragged", 200, 1000, 200, MAJ["K"], MAJ["MN"], block_n=256, cta_group=cg)
    ok &= run_case(f"cta{cg} ragged2", 392, 264, 328, MAJ["MN"], MAJ["MN"], block_n=128, cta_group=cg)
    ok &= run_case(f"cta{cg} ragged3", 520, 136, 72, MAJ["K"], MAJ["K"], block_n=128, cta_group=cg)
    ok &= run_case(f"cta{cg} direct-store", 256, 512, 256, MAJ["K"], MAJ["MN"], block_n=256, cta_group=cg, no_tma_store=1)
    return ok


def group_epilogue():
    ok = True
    gen = torch.Generator(device="cuda").manual_seed(7)
    m, n, k = 384, 520, 192
    for out_dtype in (torch.float32, torch.bfloat16):
        for cg in (1, 2):
            a, b = make_operands(m, n, k, ops.MAJOR_K, ops.MAJOR_MN, torch.bfloat16, gen)
            es = torch.randn(m, n, generator=gen, device="cuda")
            cs = torch.randn(3, n, generator=gen, device="cuda")
            cb = torch.randn(3, n, generator=gen, device="cuda")
            add = torch.randn(m, n, generator=gen, device="cuda")
            raw = torch.empty(m, n, device="cuda")
            rs = torch.zeros(m, device="cuda")
            out = ops.gemm(a, b, m, n, k, out_dtype=out_dtype, alpha=0.5, beta=2.0, act=ops.ACT_RELU, group_rows=128,
                           elem_scale=es, col_scale=cs, addend=add, col_bias=cb, raw=raw, row_sum=rs, cta_group=cg)
            torch.cuda.synchronize()
            acc = 0.5 * (a.double() @ b.double())
            g = torch.arange(m, device="cuda") // 128
            ref = torch.relu(acc * es.double() * cs.double()[g] + 2.0 * add.double() + cb.double()[g])
            e1 = (out.double() - ref).abs().max().item() / ref.abs().max().item()
            e2 = (raw.double() - acc).abs().max().item() / acc.abs().max().item()
            e3 = (rs.double() - ref.sum(1)).abs().max().item() / ref.sum(1).abs().max().item()
            tol = 1e-2 if out_dtype == torch.bfloat16 else 2e-3
            good = e1 < tol and e2 < 2e-3 and e3 < 2e-3
            ok &= good
            print(f"{'ok  ' if good else 'FAIL'} epilogue out={out_dtype} cta{cg}: out {e1:.2e} raw {e2:.2e} row_sum {e3:.2e}")
    # cos epilogue
    a, b = make_operands(256, 256, 128, ops.MAJOR_K, ops.MAJOR_K, torch.bfloat16, gen)
    cb = torch.rand(1, 256, generator=gen, device="cuda") * 6.28
    out = ops.gemm(a, b, 256, 256, 128, major_b=ops.MAJOR_K, out_dtype=torch.float32, act=ops.ACT_COS, act_scale=0.25, col_bias=cb)
    torch.cuda.synchronize()
    ref = torch.cos(a.double() @ b.double().t() + cb.double()) * 0.25
    e = (out.double() - ref).abs().max().item()
    print(f"{'ok  ' if e < 1e-4 else 'FAIL'} cos epilogue abs err {e:.2e}")
    ok &= e < 1e-4
    return ok


def group_f32():
    ok = True
    for ma in ("K", "MN"):
        for mb in ("K", "MN"):
            ok &= run_case(f"f32 A={ma} B={mb}", 200, 136, 328, MAJ[ma], MAJ[mb], dtype=torch.float32)
    return ok


def group_perf():
    for (m, n, k) in [(4096, 4096, 4096), (4096, 4096, 1024), (8192, 8192, 8192)]:
        for cg in (1, 2):
            for bn in (128, 256):
                gen = torch.Generator(device="cuda").manual_seed(1)
                a, b = make_operands(m, n, k, ops.MAJOR_K, ops.MAJOR_MN, torch.bfloat16, gen)
                out = torch.empty(m, n, dtype=torch.bfloat16, device="cuda")
                for _ in range(3):
                    ops.gemm(a, b, m, n, k, out=out, block_n=bn, cta_group=cg)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 10
                e0.record()
                for _ in range(reps):
                    ops.gemm(a, b, m, n, k, out=out, block_n=bn, cta_group=cg)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                print(f"perf m={m} n={n} k={k} cta{cg} bn={bn}: {ms:.3f} ms  {2*m*n*k/ms/1e9:.1f} TFLOP/s")
        a, b = make_operands(m, n, k, ops.MAJOR_K, ops.MAJOR_MN, torch.bfloat16, torch.Generator(device="cuda").manual_seed(1))
        for _ in range(3):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"perf m={m} n={n} k={k} cuBLAS(torch.matmul): {ms:.3f} ms  {2*m*n*k/ms/1e9:.1f} TFLOP/s")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", required=True)
    args = ap.parse_args()
    _lib.load()
    print(f"== group {args.group} on {torch.cuda.get_device_name(0)}", flush=True)
    fn = {"cta1": lambda: group_cta(1), "cta2": lambda: group_cta(2), "epilogue": group_epilogue, "f32": group_f32,
          "perf": group_perf}[args.group]
    good = fn()
    print(f"== group {args.group}: {'ALL OK' if good else 'FAILURES'}", flush=True)
    sys.exit(0 if good else 1)
