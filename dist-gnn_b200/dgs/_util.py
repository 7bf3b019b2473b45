"""Small host-side helpers shared by dgs.ops and dgs.classes."""

import torch


ID_DTYPES = {torch.int32: 0, torch.int64: 1}
# values the reference accepts (DGS_VALUE_TYPE_SWITCH, src/common/dgs_headers.h:60-74) plus the
# 16-bit floats config 5 needs; rows are moved as raw bytes so any fixed-size dtype works.
MAX_TWO_PASS_ELEMS = 1 << 26


def itype(t, what="ids"):
    try:
        return ID_DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError(f"{what} can only be int32 or int64 (got {t.dtype})") from None


def stream():
    """Raw cudaStream_t of torch's current stream on the current device."""
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def ptr(t):
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def check_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")


def check_cpu(t, name):
    if t.is_cuda:
        raise RuntimeError(f"{name} must be a CPU tensor")


def check_device_readable(t, name):
    """CUDA tensors, or CPU tensors the GPU can dereference (pinned / cudaHostRegister'ed).

    The reference dereferences the caller's CPU tensors from kernels without checking
    (src/sampling/sampler.cc:82-86) - an un-pinned tensor is an illegal address there."""
    if t.is_cuda or t.numel() == 0:
        return
    if not t.is_pinned():
        raise RuntimeError(f"{name} is a CPU tensor that is neither pinned nor registered with "
                           "_CAPI_tensor_pin_memory; kernels cannot read it")


def contiguous(t, name):
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    return t


def device_of_current():
    return torch.device("cuda", torch.cuda.current_device())


def zero_ws(nbytes, device):
    """Workspace whose leading control words are zero (the kernels leave them zero)."""
    ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
    ws[:256].zero_()
    return ws


class _Blob:
    """__cuda_array_interface__ carrier for a raw device pointer (non-owning view)."""

    def __init__(self, p, nbytes, owner=None):
        self.__cuda_array_interface__ = {
            "shape": (int(nbytes),), "typestr": "|u1", "data": (int(p), False), "version": 2,
            "strides": None,
        }
        self._owner = owner


def tensor_from_ptr(p, nbytes, dtype, shape, owner=None, device=None):
    """Non-owning tensor over [p, p + nbytes) - the counterpart of torch::from_blob in
    TensorP2PServer::Get{Local,}DeviceTensor (src/cache/tensor_p2p_cache.cc:120-132)."""
    if nbytes == 0:
        return torch.empty(shape, dtype=dtype, device=device or device_of_current())
    raw = torch.as_tensor(_Blob(p, nbytes, owner), device=device or device_of_current())
    t = raw.view(dtype).reshape(shape)
    t._dgs_owner = owner  # keep the shard alive as long as the view
    return t
