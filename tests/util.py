"""Shared helpers for the GPU parity tests."""
import torch


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max|b| — the 'max relative error' of BASELINE.md section 5."""
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).float()


def nhwc(t: torch.Tensor) -> torch.Tensor:
    """logical NCHW fp32 -> physical NHWC bf16 contiguous."""
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(t: torch.Tensor) -> torch.Tensor:
    return t.permute(0, 3, 1, 2).float()
