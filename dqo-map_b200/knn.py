"""`simple_knn._C.distCUDA2` over the C-ABI (reference submodules/simple-knn/spatial.cu:15-28)."""
import torch

from ._lib import check, lib, ptr


def distCUDA2(points):
    """points [P,3] float32 CUDA -> (mean squared distance to the 3 nearest neighbours [P], their indices [P,3] int32).
    This fork of simple-knn returns a tuple (spatial.cu:27), unlike upstream 3DGS."""
    if not points.is_cuda:
        raise RuntimeError("distCUDA2 expects a CUDA tensor (there is no CPU path)")
    P = points.size(0)
    pts = points.contiguous()
    if pts.dtype != torch.float32:
        raise TypeError("distCUDA2 expects float32 points")
    means = torch.zeros((P,), dtype=torch.float32, device=points.device)
    indices = torch.zeros((P, 3), dtype=torch.int32, device=points.device)
    if P == 0:
        return means, indices
    L = lib()
    nbytes = L.dqo_knn_workspace_bytes(P)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=points.device)
    with torch.cuda.device(points.device):
        check(L.dqo_knn3(P, ptr(pts), ptr(means), ptr(indices), ptr(ws), nbytes,
                         torch.cuda.current_stream().cuda_stream), "dqo_knn3")
    return means, indices
