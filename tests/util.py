"""helpers shared by the parity tests"""
import torch


def bf16_close(got: torch.Tensor, ref: torch.Tensor, ulps: float = 1.0, atol: float = 1e-6,
               max_mismatch_frac: float = 0.02, what: str = ""):
    """got/ref: float tensors holding bf16-representable values.  Passes when every element is
    within `ulps` bf16 ulps (taken as 2^-7 relative to max(|got|,|ref|): the spacing of bf16 at the bottom of a binade) and at most
    `max_mismatch_frac` of the elements differ at all (different fp32 accumulation order can flip
    the final rounding of a few elements, nothing more)."""
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(got).all(), f"{what}: non-finite values in result"
    diff = (got - ref).abs()
    tol = ulps * 2.0 ** -7 * torch.maximum(got.abs(), ref.abs()) + atol
    bad = diff > tol
    frac = (diff > 0).float().mean().item()
    if bad.any() or frac > max_mismatch_frac:
        idx = bad.nonzero()[:8].tolist() if bad.any() else []
        raise AssertionError(
            f"{what}: {int(bad.sum())} of {bad.numel()} elements out of tolerance, "
            f"{frac:.4%} differ at all (limit {max_mismatch_frac:.2%}); max diff {diff.max().item():.4g}; "
            f"first bad {idx}; got {[got[tuple(i)].item() for i in idx]} ref {[ref[tuple(i)].item() for i in idx]}")
    return frac
