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


def parity_json_path() -> str:
    """Where the parity tests write their MEASURED errors (copied to profiles/r02_parity.json after a GPU run)."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return os.environ.get("VCD_PARITY_JSON", os.path.join(root, "gpurun_out", "r02_parity.json"))


def record_parity(section: str, data: dict) -> None:
    """Merge {section: data} into the parity JSON (one file per test session, rewritten atomically)."""
    import json
    import os
    path = parity_json_path()
    os.makedirs(os.path.dirname(path), exist_ok=True)
    try:
        with open(path) as f:
            cur = json.load(f)
    except Exception:
        cur = {}
    cur[section] = data
    tmp = path + ".tmp"
    with open(tmp, "w") as f:
        json.dump(cur, f, indent=1, sort_keys=True)
    os.replace(tmp, path)
    print(f"[parity] {section}: " + json.dumps(data, sort_keys=True))
