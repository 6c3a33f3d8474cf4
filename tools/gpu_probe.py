"""GPU-side diagnostics + microbenchmarks (development tool, run under gpurun).

  python tools/gpu_probe.py gemm_check   # correctness of the tcgen05 GEMM with an error map
  python tools/gpu_probe.py bench        # per-kernel throughput vs roofline / cuBLAS bar
Writes JSON lines to gpurun_out/probe_<mode>.jsonl
"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from llmq_b200 import lib

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
BF = torch.bfloat16
dev = torch.device("cuda", 0)


def emit(f, **kw):
    f.write(json.dumps(kw) + "\n")
    f.flush()
    print(json.dumps(kw))


def timeit(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def gemm_check(f):
    torch.manual_seed(0)
    mode = int(os.environ.get("GEMM_MODE", "1"))
    lib.gemm_set_mode(mode)
    for (M, N, K) in [(128, 64, 64), (256, 256, 64), (256, 256, 128), (256, 256, 256), (256, 512, 1024), (200, 4096, 4096), (1000, 6144, 4096), (4096, 4096, 14336)]:
        for bn in (64, 128, 256):
            if N % bn or (mode == 2 and bn == 64):
                continue
            a = torch.randn(M, K, device=dev).to(BF)
            w = (torch.randn(N, K, device=dev) * 0.05).to(BF)
            c = torch.full((M, N), float("nan"), dtype=BF, device=dev)
            lib.gemm_set_tile_n(bn)
            lib.gemm_bf16(a, w, c)
            torch.cuda.synchronize()
            ref = (a.float() @ w.float().t())
            d = (c.float() - ref).abs()
            nan = torch.isnan(c.float()).sum().item()
            rel = (d / (ref.abs() + 1e-3))
            bad = (d > 0.02 * ref.abs() + 0.01)
            info = dict(kind="gemm_check", mode=mode, M=M, N=N, K=K, bn=bn, nan=nan, max_abs=float(d[~torch.isnan(d)].max().item()) if nan < d.numel() else None,
                        bad_frac=float(bad.float().mean().item()))
            if bad.any():
                # error map at 32x32 granularity (fraction of bad elements per cell)
                m32, n32 = (M + 31) // 32, (N + 31) // 32
                pad = torch.zeros(m32 * 32, n32 * 32, device=dev)
                pad[:M, :N] = bad.float()
                cell = pad.view(m32, 32, n32, 32).mean((1, 3))
                info["errmap_rows"] = [[round(x, 2) for x in r[:16]] for r in cell[:8].tolist()]
                # does the result match a reference computed from a K-permuted / partial-K input?
                for kk in (16, 32, 64):
                    part = a[:, :kk].float() @ w[:, :kk].float().t()
                    info[f"matches_first_{kk}_k"] = float(((c.float() - part).abs() < 0.02 * part.abs() + 0.01).float().mean().item())
                info["sample_got"] = c[:2, :4].float().tolist()
                info["sample_ref"] = ref[:2, :4].tolist()
            emit(f, **info)
    lib.gemm_set_tile_n(0)
    lib.gemm_set_mode(0)


def gemm2_bench(f):
    """1-CTA vs 2-CTA (cta_group::2) vs cuBLAS on the Llama-3-8B layer shapes"""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    shapes = {"qkv": (6144, 4096), "o": (4096, 4096), "gate_up": (28672, 4096), "down": (4096, 14336), "lm_head": (128256, 4096)}
    for M in (256, 512, 1024, 2048, 4096, 8192):
        for name, (N, K) in shapes.items():
            if name == "lm_head" and M > 1024:
                continue
            a = torch.randn(M, K, device=dev).to(BF)
            w = (torch.randn(N, K, device=dev) * 0.02).to(BF)
            c = torch.empty(M, N, dtype=BF, device=dev)
            res = dict(kind="gemm2", name=name, M=M, N=N, K=K)
            for mode, bn in ((1, 256), (1, 128), (2, 256), (2, 128)):
                lib.gemm_set_mode(mode)
                lib.gemm_set_tile_n(bn)
                med, _ = timeit(lambda: lib.gemm_bf16(a, w, c), iters=10, flush=flush)
                res[f"ms_m{mode}_bn{bn}"] = round(med, 4)
            lib.gemm_set_mode(0)
            lib.gemm_set_tile_n(0)
            med, _ = timeit(lambda: lib.gemm_bf16(a, w, c), iters=10, flush=flush)
            res["ms_auto"] = round(med, 4)
            medc, _ = timeit(lambda: torch.matmul(a, w.t(), out=c), iters=10, flush=flush)
            res["ms_cublas"] = round(medc, 4)
            fl = 2.0 * M * N * K
            for k in list(res):
                if k.startswith("ms_"):
                    res["tf_" + k[3:]] = round(fl / res[k] / 1e9, 1)
            res["resident_pairs"] = lib.load().b200q_gemm_resident_pairs()
            emit(f, **res)
            del a, w, c


def ncu_targets(f):
    """the bench's dominant launches in isolation (small process => cheap to put under ncu):
    the four projection GEMMs at the decode batch M=4608 and the decode attention at B=4608"""
    M = int(os.environ.get("NCU_M", "4608"))
    shapes = {"qkv": (6144, 4096), "o": (4096, 4096), "down": (4096, 14336)}
    for name, (N, K) in shapes.items():
        a = torch.randn(M, K, device=dev).to(BF)
        w = (torch.randn(N, K, device=dev) * 0.02).to(BF)
        c = torch.empty(M, N, dtype=BF, device=dev)
        for _ in range(2):
            lib.gemm_bf16(a, w, c)
        torch.cuda.synchronize()
        del a, w, c
    a = torch.randn(M, 4096, device=dev).to(BF)
    w = (torch.randn(28672, 4096, device=dev) * 0.02).to(BF)
    c = torch.empty(M, 14336, dtype=BF, device=dev)
    for _ in range(2):
        lib.gemm_swiglu_bf16(a, w, c)
    torch.cuda.synchronize()
    del a, w, c
    n_q, n_kv, D, BS, ctx = 32, 8, 128, 16, 192
    nb_per = ctx // BS
    NB = M * nb_per + 1
    kv = torch.randn(NB, 2, n_kv, BS, D, device=dev).to(BF)
    bt = (torch.randperm(NB - 1, device=dev).to(torch.int32).view(M, nb_per) + 1)
    bt = torch.cat([bt, torch.zeros(M, (-nb_per) % 8, dtype=torch.int32, device=dev)], 1).contiguous()
    ctxs = torch.full((M,), ctx, dtype=torch.int32, device=dev)
    qkv = torch.randn(M, (n_q + 2 * n_kv) * D, device=dev).to(BF)
    out = torch.empty(M, n_q * D, dtype=BF, device=dev)
    for _ in range(2):
        lib.decode_attn(qkv, out, kv, bt, ctxs, n_q, n_kv, D, BS, 1 / math.sqrt(D))
    torch.cuda.synchronize()
    emit(f, kind="ncu_targets", M=M, done=True)


def decode_variants(f):
    """split (v1) vs streaming (v2) decode attention on the bench's shape: B=4608, contexts 129..255"""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n_q, n_kv, D, BS = 32, 8, 128, 16
    L = lib.load()
    for B, lo, hi in ((4608, 129, 256), (4608, 192, 193), (1024, 129, 256), (296, 129, 256), (64, 1024, 1025)):
        g = torch.Generator(device="cpu").manual_seed(0)
        ctx_list = torch.randint(lo, hi, (B,), generator=g)
        nb = ((ctx_list + BS - 1) // BS)
        maxb = int(nb.max())
        NB = int(nb.sum()) + 1
        kv = torch.randn(NB, 2, n_kv, BS, D, device=dev).to(BF)
        perm = torch.randperm(NB - 1) + 1
        bt = torch.zeros(B, (maxb + 7) // 8 * 8, dtype=torch.int32)
        off = 0
        for i in range(B):
            n = int(nb[i])
            bt[i, :n] = perm[off:off + n]
            off += n
        bt = bt.to(dev)
        ctxs = ctx_list.to(torch.int32).to(dev)
        qkv = torch.randn(B, (n_q + 2 * n_kv) * D, device=dev).to(BF)
        out = torch.empty(B, n_q * D, dtype=BF, device=dev)
        byts = float(ctx_list.sum()) * n_kv * D * 2 * 2
        res = dict(kind="decode_variants", B=B, ctx_lo=lo, ctx_hi=hi - 1)
        for v in (1, 2):
            L.b200q_decode_attn_set_variant(v)
            med, best = timeit(lambda: lib.decode_attn(qkv, out, kv, bt, ctxs, n_q, n_kv, D, BS, 1 / math.sqrt(D)), iters=15, flush=flush)
            res[f"ms_v{v}"] = round(med, 4)
            res[f"gbs_v{v}"] = round(byts / med / 1e6, 1)
        L.b200q_decode_attn_set_variant(0)
        emit(f, **res)
        del kv


def gemm_small_m(f):
    """decode-sized M: split-K (auto) vs no split vs cuBLAS; weight stream GB/s is the figure of merit"""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    shapes = {"qkv": (6144, 4096), "o": (4096, 4096), "down": (4096, 14336)}
    L = lib.load()
    for M in (16, 64, 128, 256):
        for name, (N, K) in shapes.items():
            a = torch.randn(M, K, device=dev).to(BF)
            w = (torch.randn(N, K, device=dev) * 0.02).to(BF)
            c = torch.empty(M, N, dtype=BF, device=dev)
            res = dict(kind="gemm_small_m", name=name, M=M, N=N, K=K)
            for tag, sk in (("auto", 0), ("nosplit", 1), ("s2", 2), ("s4", 4), ("s8", 8)):
                if sk > 1 and (K // 64) % sk:
                    continue
                L.b200q_gemm_set_splitk(sk)
                med, _ = timeit(lambda: lib.gemm_bf16(a, w, c), iters=15, flush=flush)
                res[f"us_{tag}"] = round(med * 1e3, 2)
            L.b200q_gemm_set_splitk(0)
            medc, _ = timeit(lambda: torch.matmul(a, w.t(), out=c), iters=15, flush=flush)
            res["us_cublas"] = round(medc * 1e3, 2)
            res["weight_gbs_auto"] = round(N * K * 2 / res["us_auto"] / 1e3, 1)
            res["weight_gbs_cublas"] = round(N * K * 2 / res["us_cublas"] / 1e3, 1)
            emit(f, **res)
            del a, w, c


def ncu_small_m(f):
    """decode-sized GEMMs in isolation for an `ncu --set full` capture: fused gate_up at M = 16 and 128,
    split-K qkv / down at M = 128"""
    for M in (16, 128):
        a = torch.randn(M, 4096, device=dev).to(BF)
        w = (torch.randn(28672, 4096, device=dev) * 0.02).to(BF)
        c = torch.empty(M, 14336, dtype=BF, device=dev)
        for _ in range(2):
            lib.gemm_swiglu_bf16(a, w, c)
        torch.cuda.synchronize()
        del a, w, c
    M = 128
    for N, K in ((6144, 4096), (4096, 14336)):
        a = torch.randn(M, K, device=dev).to(BF)
        w = (torch.randn(N, K, device=dev) * 0.02).to(BF)
        c = torch.empty(M, N, dtype=BF, device=dev)
        for _ in range(2):
            lib.gemm_bf16(a, w, c)
        torch.cuda.synchronize()
        del a, w, c
    emit(f, kind="ncu_small_m", done=True)


def stream_probe(f):
    """how fast can one CTA per SM pull a weight matrix through a shared-memory ring?  2-D TMA boxes in
    the GEMM's own pattern (rows x 128 B, rows K*2 bytes apart) vs 1-D bulk copies of the same bytes laid
    out contiguously (a pre-tiled weight layout) — for the fused gate_up shape [28672, 4096] and qkv"""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    L = lib.load()
    for name, (N, K) in {"gate_up": (28672, 4096), "qkv": (6144, 4096), "down": (4096, 14336)}.items():
        w = (torch.randn(N, K, device=dev) * 0.02).to(BF)
        for rows, stages in ((256, 4), (256, 6), (128, 8), (128, 12), (64, 16), (64, 24)):
            for grid in (148, N // rows if N // rows < 148 else 148):
                for mode in (0, 1):
                    fn = lambda: lib.check(L.b200q_stream_probe(w.data_ptr(), N, K, rows, stages, mode, grid, 0))
                    torch.cuda.synchronize()
                    med, best = timeit(lambda: (fn(), None)[1], iters=11, flush=flush)
                    emit(f, kind="stream_probe", name=name, N=N, K=K, rows=rows, stages=stages, ring_kb=rows * stages // 8,
                         grid=grid, mode="tma2d" if mode == 0 else "bulk1d", us=round(med * 1e3, 2),
                         gbs=round(N * K * 2 / med / 1e6, 1), gbs_per_cta=round(N * K * 2 / med / 1e6 / grid, 1))
        del w


def gemm_limits(f):
    """where do the GEMM's bubbles come from?  time the kernel with the operand loads and/or the
    epilogue switched off (results are garbage in those modes; timing only)"""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    shapes = {"qkv": (6144, 4096), "down": (4096, 14336), "gate_up": (28672, 4096)}
    L = lib.load()
    for M in (4608, 9472):
        for name, (N, K) in shapes.items():
            a = torch.randn(M, K, device=dev).to(BF)
            w = (torch.randn(N, K, device=dev) * 0.02).to(BF)
            c = torch.empty(M, N, dtype=BF, device=dev)
            res = dict(kind="gemm_limits", name=name, M=M, N=N, K=K)
            fl = 2.0 * M * N * K
            for mode in (1, 2):
                for dbg in (0, 1, 2, 3):
                    lib.gemm_set_mode(mode)
                    lib.gemm_set_tile_n(256)
                    L.b200q_gemm_set_debug(dbg)
                    med, _ = timeit(lambda: lib.gemm_bf16(a, w, c), iters=8, flush=flush)
                    res[f"tf_cta{mode}_dbg{dbg}"] = round(fl / med / 1e9, 1)
            L.b200q_gemm_set_debug(0)
            lib.gemm_set_mode(0)
            lib.gemm_set_tile_n(0)
            medc, _ = timeit(lambda: torch.matmul(a, w.t(), out=c), iters=8, flush=flush)
            res["tf_cublas"] = round(fl / medc / 1e9, 1)
            emit(f, **res)
            del a, w, c


def argmax_ties(f):
    """which index does torch.argmax return on exact ties on this GPU? (vLLM's greedy sampler is
    logits.argmax(-1): vllm/v1/sample/sampler.py:235-236)"""
    for V in (2048, 128256):
        for dt in (torch.float32, torch.bfloat16):
            x = torch.randn(64, V, device=dev).to(dt)
            big = x.max().item() + 1.0
            idx = torch.randint(0, V, (64, 3), device=dev)
            x.scatter_(1, idx, big)
            got = x.argmax(-1).cpu()
            lo, hi = idx.min(1).values.cpu(), idx.max(1).values.cpu()
            emit(f, kind="argmax_ties", V=V, dtype=str(dt), picks_lowest=int((got == lo).sum()),
                 picks_highest=int((got == hi).sum()), rows=64)
            ours = torch.empty(64, dtype=torch.int32, device=dev)
            if dt == torch.bfloat16:
                lib.argmax_bf16(x, ours)
                emit(f, kind="argmax_ties_ours", V=V, picks_lowest=int((ours.cpu() == lo).sum()), rows=64)


def bench(f):
    peaks = {"hbm_gbs": 6480.5, "bf16_tflops": 1707.6}
    try:
        peaks.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
    except Exception:
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    # ---- GEMM: ours vs cuBLAS (torch.matmul) on the Llama-3-8B layer shapes
    shapes = {"qkv": (6144, 4096), "o": (4096, 4096), "gate_up": (28672, 4096), "down": (4096, 14336), "lm_head": (128256, 4096)}
    for M in (64, 256, 1024, 2048, 4096):
        for name, (N, K) in shapes.items():
            if name == "lm_head" and M > 1024:
                continue
            a = torch.randn(M, K, device=dev).to(BF)
            w = (torch.randn(N, K, device=dev) * 0.02).to(BF)
            c = torch.empty(M, N, dtype=BF, device=dev)
            res = dict(kind="gemm", name=name, M=M, N=N, K=K)
            for bn in (0, 64, 128, 256):
                lib.gemm_set_tile_n(bn)
                med, best = timeit(lambda: lib.gemm_bf16(a, w, c), iters=10, flush=flush)
                res[f"ms_bn{bn}"] = round(med, 4)
            lib.gemm_set_tile_n(0)
            medc, bestc = timeit(lambda: torch.matmul(a, w.t(), out=c), iters=10, flush=flush)
            res["ms_cublas"] = round(medc, 4)
            fl = 2.0 * M * N * K
            best_ours = min(res[f"ms_bn{b}"] for b in (64, 128, 256))
            res["tflops_ours_best"] = round(fl / best_ours / 1e9, 1)
            res["tflops_auto"] = round(fl / res["ms_bn0"] / 1e9, 1)
            res["tflops_cublas"] = round(fl / medc / 1e9, 1)
            res["weight_gbs_auto"] = round(N * K * 2 / res["ms_bn0"] / 1e6, 1)
            emit(f, **res)
            del a, w, c
    # ---- decode attention: algorithmic bytes = sum ctx * n_kv * D * 2 (K,V) * 2 B per layer
    n_q, n_kv, D, BS = 32, 8, 128, 16
    for B, ctx in ((64, 192), (256, 192), (1024, 192), (256, 1024), (16, 2048)):
        nb_per = (ctx + BS - 1) // BS
        NB = B * nb_per + 1
        kv = torch.randn(NB, 2, n_kv, BS, D, device=dev).to(BF)
        bt = torch.randperm(NB - 1, device=dev).to(torch.int32).view(B, nb_per) + 1
        pad = (-nb_per) % 8
        if pad:
            bt = torch.cat([bt, torch.zeros(B, pad, dtype=torch.int32, device=dev)], 1).contiguous()
        ctxs = torch.full((B,), ctx, dtype=torch.int32, device=dev)
        qkv = torch.randn(B, (n_q + 2 * n_kv) * D, device=dev).to(BF)
        out = torch.empty(B, n_q * D, dtype=BF, device=dev)
        med, best = timeit(lambda: lib.decode_attn(qkv, out, kv, bt, ctxs, n_q, n_kv, D, BS, 1 / math.sqrt(D)), iters=20, flush=flush)
        byts = B * ctx * n_kv * D * 2 * 2
        emit(f, kind="decode_attn", B=B, ctx=ctx, ms=round(med, 4), ms_best=round(best, 4), gbs=round(byts / med / 1e6, 1),
             frac_of_measured_hbm=round(byts / med / 1e6 / peaks["hbm_gbs"], 3))
        del kv
    # ---- prefill attention 128-token prompts
    for nseq in (16, 64):
        P = 128
        nb_per = P // BS
        NB = nseq * nb_per + 1
        kv = torch.randn(NB, 2, n_kv, BS, D, device=dev).to(BF)
        bt = (torch.arange(nseq * nb_per, device=dev, dtype=torch.int32).view(nseq, nb_per) + 1).contiguous()
        T = nseq * P
        qkv = torch.randn(T, (n_q + 2 * n_kv) * D, device=dev).to(BF)
        out = torch.empty(T, n_q * D, dtype=BF, device=dev)
        tiles = torch.tensor([[s, s * P + j, 16, j] for s in range(nseq) for j in range(0, P, 16)], dtype=torch.int32, device=dev)
        med, best = timeit(lambda: lib.prefill_attn(qkv, out, kv, bt, tiles, n_q, n_kv, D, BS, 1 / math.sqrt(D)), iters=10, flush=flush)
        fl = 4.0 * nseq * n_q * D * P * P / 2
        emit(f, kind="prefill_attn", nseq=nseq, P=P, ms=round(med, 4), tflops_causal=round(fl / med / 1e9, 1))
    # ---- elementwise kernels (bytes = reads + writes)
    T, H, I = 2048, 4096, 14336
    x = torch.randn(T, H, device=dev).to(BF)
    r = torch.randn(T, H, device=dev).to(BF)
    wn = torch.ones(H, device=dev, dtype=BF)
    med, _ = timeit(lambda: lib.add_rmsnorm(x, r, wn, 1e-5), flush=flush)
    emit(f, kind="add_rmsnorm", T=T, H=H, ms=round(med, 4), gbs=round(4 * T * H * 2 / med / 1e6, 1))
    gu = torch.randn(T, 2 * I, device=dev).to(BF)
    act = torch.empty(T, I, dtype=BF, device=dev)
    med, _ = timeit(lambda: lib.swiglu(gu, act), flush=flush)
    emit(f, kind="swiglu", T=T, I=I, ms=round(med, 4), gbs=round(3 * T * I * 2 / med / 1e6, 1))
    qkv = torch.randn(T, 6144, device=dev).to(BF)
    table = torch.randn(4096, 128, device=dev).to(BF)
    pos = torch.randint(0, 4096, (T,), dtype=torch.int32, device=dev)
    slots = torch.randperm(4096, device=dev)[:T].to(torch.int32)
    kvl = torch.zeros(256, 2, 8, 16, 128, dtype=BF, device=dev)
    med, _ = timeit(lambda: lib.rope_kvwrite(qkv, table, pos, slots, kvl, 32, 8, 128, 16), flush=flush)
    emit(f, kind="rope_kvwrite", T=T, ms=round(med, 4), gbs=round((T * 6144 * 2 + T * 4096 * 2 + T * 2048 * 2) / med / 1e6, 1))
    logits = torch.randn(1024, 128256, device=dev).to(BF)
    ids = torch.empty(1024, dtype=torch.int32, device=dev)
    med, _ = timeit(lambda: lib.argmax_bf16(logits, ids), flush=flush)
    emit(f, kind="argmax", B=1024, V=128256, ms=round(med, 4), gbs=round(1024 * 128256 * 2 / med / 1e6, 1))


if __name__ == "__main__":
    mode = sys.argv[1]
    lib.require_device()
    tag = os.environ.get("PROBE_TAG", "")
    with open(os.path.join(OUT, f"probe_{mode}{tag}.jsonl"), "w") as f:
        {"gemm_check": gemm_check, "bench": bench, "gemm2_bench": gemm2_bench, "ncu_targets": ncu_targets,
         "argmax_ties": argmax_ties, "gemm_limits": gemm_limits, "decode_variants": decode_variants, "gemm_small_m": gemm_small_m,
         "stream_probe": stream_probe, "ncu_small_m": ncu_small_m}[mode](f)
