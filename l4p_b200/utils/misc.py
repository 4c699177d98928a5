"""Mirror of l4p/utils/misc.py (apply_fn :11-38, safe_inverse :48-62) for device tensors."""
import torch


def apply_fn(x: torch.Tensor, fn_type: str = "linear") -> torch.Tensor:
    if fn_type == "log":
        out = torch.log(x)
    elif fn_type == "exp":
        out = torch.exp(x)
    elif fn_type == "sigmoid":
        out = torch.sigmoid(x)
    elif fn_type == "linear":
        out = x
    elif fn_type == "inverse":
        out = torch.where(x.abs() > 1e-8, 1.0 / x, torch.zeros_like(x))
    else:
        print(f"Not implemented {fn_type}")
        raise NotImplementedError
    return out.to(x.dtype)


def safe_inverse(depth_or_disp: torch.Tensor, keep_above: float = 0.0) -> torch.Tensor:
    assert isinstance(depth_or_disp, torch.Tensor)
    return torch.where(depth_or_disp > keep_above, 1.0 / depth_or_disp, torch.zeros_like(depth_or_disp))
