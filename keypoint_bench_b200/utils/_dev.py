"""Device plumbing shared by the drop-in wrappers: the reference's functions are device-agnostic
torch code, so the stand-ins accept tensors on any device, run on the current CUDA device and hand
results back on the caller's device.  There is no CPU compute path."""
import torch

from .. import _lib


def to_cuda(t: torch.Tensor) -> torch.Tensor:
    if t.is_cuda:
        return t
    if not torch.cuda.is_available():
        raise _lib.KbError('keypoint_bench_b200 needs a CUDA device (no CPU fallback)')
    return t.cuda(non_blocking=True)


def like(result: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    return result if result.device == ref.device else result.to(ref.device)


def as_int(v) -> int:
    return int(v.item()) if isinstance(v, torch.Tensor) else int(v)
