"""Closed-form synthetic graphs / features of the shapes BASELINE.json names.

Everything is a pure integer function of (seed, index) built from a 64-bit mixer, so any rank can
regenerate any node's row, degree or feature vector without a host copy, CPU and GPU generation
agree bit for bit, and shards can be generated directly on their owner GPU (papers100M- and
friendster-sized inputs never exist on the host).

  degrees   : power law (P(deg ~ d) ~ d^-2) from the number of leading zeros of a hash, clipped to
              `classes` octaves, ~1/16 of the nodes isolated (so deg = 0, deg <= k and deg > k paths
              are all exercised)
  indices   : indices[e] = mix(seed, e) mod N
  features  : feat[n, j] = (24-bit hash - 2^23) / 2^20   (exact in fp32; /2^4 of an int8 for bf16)
  weights   : w[e] = (1 + 10-bit hash) / 256  in (0, 4]   (positive, exact in fp32)
"""
import math

import torch

SHAPES = {
    # name: (nodes, target edges, feature dim, feature dtype)
    "tiny": (1000, 20_000, 16, torch.float32),
    "small": (100_000, 2_500_000, 100, torch.float32),
    "products": (2_449_029, 61_859_140, 100, torch.float32),
    "papers100M": (111_059_956, 1_615_685_872, 128, torch.float32),
    "friendster": (65_608_366, 1_806_067_135, 256, torch.bfloat16),
}

_M1 = -49064778989728563      # 0xff51afd7ed558ccd as int64
_M2 = -4265267296055464877    # 0xc4ceb9fe1a85ec53 as int64
_GOLD = -7046029254386353131  # 0x9e3779b97f4a7c15 as int64


def _lsr(x, s):
    """logical shift right on int64 tensors"""
    return (x >> s) & ((1 << (64 - s)) - 1)


def mix64(x):
    """murmur3 fmix64 on int64 tensors (wrapping arithmetic, identical on CPU and CUDA)."""
    x = x ^ _lsr(x, 33)
    x = x * _M1
    x = x ^ _lsr(x, 33)
    x = x * _M2
    x = x ^ _lsr(x, 33)
    return x


def _wrap(x):
    """python int -> two's-complement int64"""
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


def _hash(seed, idx):
    return mix64(idx + _wrap((seed + 1) * _GOLD))


def degrees(num_nodes, target_edges, seed=0, classes=13, device="cpu", start=0, count=None):
    """int64 degrees of nodes [start, start + count)."""
    count = num_nodes - start if count is None else count
    # mean of the construction = 0.75 * a * (classes + 1) * 15/16  ->  solve for a (16.16 fixed)
    a = target_edges / num_nodes / (0.75 * (classes + 1) * 15.0 / 16.0)
    A = int(round(a * 65536))
    n = torch.arange(start, start + count, dtype=torch.int64, device=device)
    h = _hash(seed * 4 + 0, n)
    h32 = _lsr(h, 32)
    c = torch.zeros_like(n)
    for b in range(1, classes):
        c += (h32 < (1 << (32 - b))).to(torch.int64)
    f16 = h & 0xFFFF
    deg = ((A * (65536 + f16)) << c) >> 32
    isolated = (_lsr(h, 16) & 15) == 0
    return torch.where(isolated, torch.zeros_like(deg), deg)


def edge_targets(num_nodes, seed, start, count, device="cpu"):
    """indices[start : start + count] (int64)."""
    e = torch.arange(start, start + count, dtype=torch.int64, device=device)
    return (_hash(seed * 4 + 1, e) & 0x7FFFFFFFFFFFFFFF) % num_nodes


def edge_weights(seed, start, count, device="cpu"):
    e = torch.arange(start, start + count, dtype=torch.int64, device=device)
    return ((_hash(seed * 4 + 2, e) & 1023) + 1).to(torch.float32) / 256.0


def feature_rows(nids, dim, dtype=torch.float32, seed=0):
    """Feature rows of the given node ids (any device)."""
    j = torch.arange(dim, dtype=torch.int64, device=nids.device)
    h = _hash(seed * 4 + 3, nids.to(torch.int64)[:, None] * dim + j[None, :])
    if dtype == torch.float32:
        return ((_lsr(h, 40) - (1 << 23)).to(torch.float32) / float(1 << 20)).contiguous()
    if dtype in (torch.bfloat16, torch.float16):
        return ((_lsr(h, 56) - 128).to(torch.float32) / 16.0).to(dtype).contiguous()
    if dtype in (torch.int32, torch.int64):
        return _lsr(h, 40).to(dtype).contiguous()
    raise ValueError(f"unsupported feature dtype {dtype}")


def make_csr(num_nodes, target_edges, seed=0, device="cpu", id_dtype=torch.int64, weights=False,
             chunk=1 << 26, classes=13):
    """Full CSR on one device: (indptr int64[N+1], indices id_dtype[E], probs fp32[E] or None)."""
    deg = degrees(num_nodes, target_edges, seed, classes=classes, device=device)
    indptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=device)
    torch.cumsum(deg, 0, out=indptr[1:])
    del deg
    E = int(indptr[-1].item())
    indices = torch.empty(E, dtype=id_dtype, device=device)
    probs = torch.empty(E, dtype=torch.float32, device=device) if weights else None
    for s in range(0, E, chunk):
        c = min(chunk, E - s)
        indices[s:s + c] = edge_targets(num_nodes, seed, s, c, device).to(id_dtype)
        if weights:
            probs[s:s + c] = edge_weights(seed, s, c, device)
    return indptr, indices, probs


def make_shard(num_nodes, target_edges, rank, world, seed=0, device="cpu", id_dtype=torch.int64,
               weights=False, node_chunk=1 << 21, classes=13):
    """The CSR rows of the nodes GPU `rank` owns under modulo sharding (n % world == rank), generated
    directly on `device`: (nids, sub_indptr int64[C+1], sub_indices id_dtype[sum deg], sub_probs or
    None).  Bit-identical to slicing make_csr()'s output, without ever materialising the full edge
    list (the global indptr - 8 bytes per node - is the only whole-graph array built)."""
    deg = degrees(num_nodes, target_edges, seed, classes=classes, device=device)
    gptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=device)
    torch.cumsum(deg, 0, out=gptr[1:])
    nids = torch.arange(rank, num_nodes, world, dtype=torch.int64, device=device)
    d = deg[nids]
    del deg
    sub_indptr = torch.zeros(nids.numel() + 1, dtype=torch.int64, device=device)
    torch.cumsum(d, 0, out=sub_indptr[1:])
    starts = gptr[nids]
    del gptr
    total = int(sub_indptr[-1].item())
    sub_indices = torch.empty(total, dtype=id_dtype, device=device)
    sub_probs = torch.empty(total, dtype=torch.float32, device=device) if weights else None
    for a in range(0, nids.numel(), node_chunk):
        b = min(nids.numel(), a + node_chunk)
        lo, hi = int(sub_indptr[a].item()), int(sub_indptr[b].item())
        if hi == lo:
            continue
        # global edge id of every local edge in [lo, hi)
        shift = torch.repeat_interleave(starts[a:b] - sub_indptr[a:b], d[a:b])
        e = shift + torch.arange(lo, hi, dtype=torch.int64, device=device)
        sub_indices[lo:hi] = ((_hash(seed * 4 + 1, e) & 0x7FFFFFFFFFFFFFFF) % num_nodes).to(id_dtype)
        if weights:
            sub_probs[lo:hi] = ((_hash(seed * 4 + 2, e) & 1023) + 1).to(torch.float32) / 256.0
        del shift, e
    return nids.to(id_dtype), sub_indptr, sub_indices, sub_probs


def make_features(num_nodes, dim, dtype=torch.float32, seed=0, device="cpu", chunk=1 << 20,
                  nids=None):
    """Feature table of all nodes (or of `nids`, e.g. a shard) generated chunk by chunk."""
    if nids is None:
        n = num_nodes
        out = torch.empty((n, dim), dtype=dtype, device=device)
        for s in range(0, n, chunk):
            c = min(chunk, n - s)
            ids = torch.arange(s, s + c, dtype=torch.int64, device=device)
            out[s:s + c] = feature_rows(ids, dim, dtype, seed)
        return out
    out = torch.empty((nids.numel(), dim), dtype=dtype, device=device)
    for s in range(0, nids.numel(), chunk):
        out[s:s + chunk] = feature_rows(nids[s:s + chunk].to(device), dim, dtype, seed)
    return out


def seed_batches(num_nodes, batch, num_batches, seed=0, device="cpu", id_dtype=torch.int64):
    """`num_batches` batches of `batch` distinct node ids (a seeded permutation cut into batches,
    like SeedGenerator(shuffle=True) over the training set)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1234567 + seed)
    need = batch * num_batches
    if need <= num_nodes:
        perm = torch.randperm(num_nodes, generator=g)[:need]
    else:
        reps = math.ceil(need / num_nodes)
        perm = torch.cat([torch.randperm(num_nodes, generator=g) for _ in range(reps)])[:need]
    return perm.to(id_dtype).reshape(num_batches, batch).to(device)
