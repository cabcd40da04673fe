"""SE(3) interpolation helpers with the signatures of starster/utils.py:13,57 (host-side, tiny; not a hot path)."""
import torch

__all__ = ("interp_se3", "interp_se3_path")


def interp_se3(mat1: torch.Tensor, mat2: torch.Tensor, fac: float) -> torch.Tensor:
    """Blend translation and rotation linearly, then re-orthonormalise the rotation columns (Gram-Schmidt)."""
    out = torch.zeros_like(mat1)
    out[3, 3] = 1
    out[:3, 3] = mat1[:3, 3] + (mat2[:3, 3] - mat1[:3, 3]) * fac
    r = mat1[:3, :3] + (mat2[:3, :3] - mat1[:3, :3]) * fac
    c0, c1, c2 = r[:, 0].clone(), r[:, 1].clone(), r[:, 2].clone()
    c1 = c1 - c0 * c0.dot(c1)
    c2 = c2 - c0 * c0.dot(c2)
    c2 = c2 - c1 * c1.dot(c2)
    r = torch.stack([c0, c1, c2], dim=1)
    out[:3, :3] = r / torch.linalg.norm(r, dim=0)
    return out


def interp_se3_path(mat1: torch.Tensor, mat2: torch.Tensor, steps: int) -> torch.Tensor:
    return torch.stack([interp_se3(mat1, mat2, f) for f in torch.linspace(0, 1, steps)], dim=0)
