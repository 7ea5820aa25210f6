"""Drop-in replacements for artensor's numerical executors.

Same names, arguments and container semantics as the reference:

  tensor_contraction(tensors, scheme)                       artensor/contraction.py:62-76
  tensor_contraction_sparse(tensors, scheme, scientific_notation=False)
                                                            artensor/contraction.py:132-205

`tensors` is a list or dict of CUDA torch tensors indexed by the ids in the scheme; `scheme` is
exactly what `contraction_scheme` / `contraction_scheme_sparse` return.  Like the reference the
functions mutate the container: `tensors[i]` of the last step holds the result, consumed right
operands of the sparse executor are set to `[]` (contraction.py:174,188,191).  Intermediate
`tensors[i]` values of earlier steps are NOT materialised in the container (they live in the
device arena); that is the one container-visible difference.

Errors: the reference prints the step and calls sys.exit(1) (contraction.py:71-74, :192-195);
here a SchemeError / NativeError (both RuntimeError/ValueError subclasses) is raised.

There is no CPU path: tensors must live on a CUDA device and libtnc_b200.so must be built.
"""
import torch

from . import _native as N
from .backend import ContractionPlan, PlanOptions
from .plan import SchemeError

# torch dtype -> compute mode.  complex64: fp32-accurate split-precision tensor-core products.
# complex32 (the reference would run torch.einsum in complex-half): the reduced-precision
# tensor-core mode -- fp16 operands in the GEMM steps, fp32 accumulation, intermediates and the
# returned amplitudes stay complex64 (n53 amplitudes are ~1e-8, below the fp16 range).
_DTYPES = {torch.complex64: "c64", torch.complex32: "chalf"}


def mode_options(mode, options=None):
    """Plan options for a compute mode: "chalf" forces the single-product fp16 precision."""
    from dataclasses import replace
    options = options or default_options()
    return replace(options, tc_precision="f16") if mode == "chalf" else options

# Compiled plans, keyed by the CONTENT of the scheme (a scheme list mutated in place must not hit
# the plan of its old contents), the leaf shapes, the slicing and the options.
_PLAN_CACHE = {}
_PLAN_CACHE_MAX = 16


def scheme_fingerprint(scheme):
    """Digest of everything in a scheme that the plan depends on: edges, einsum strings, the row
    index tensors of sparse steps, reshape / next shapes."""
    import hashlib
    import numpy as np
    h = hashlib.blake2b(digest_size=16)
    for step in scheme:
        h.update(repr((tuple(step[0]), step[1], len(step))).encode())
        if len(step) >= 3:
            for side in step[2]:
                h.update(b"|")
                for idx in side:
                    arr = idx.numpy() if torch.is_tensor(idx) else np.asarray(idx)
                    h.update(np.ascontiguousarray(arr, dtype=np.int64).tobytes())
                    h.update(b";")
        if len(step) >= 5:
            h.update(repr((step[3], tuple(step[4]))).encode())
    return h.hexdigest()


def _leaf_shapes(tensors, ids):
    return {i: tuple(tensors[i].shape) for i in ids}


def _scheme_ids(scheme):
    ids = []
    seen = set()
    for s in scheme:
        for t in s[0]:
            if t not in seen:
                seen.add(t)
                ids.append(t)
    return ids


def default_options():
    return PlanOptions()


def get_plan(scheme, tensors, sparse, *, slicing_bonds=(), slicing_indices=None, dtype="c64", options=None):
    """Compile (or fetch from the cache) the plan of `scheme` for these leaf shapes."""
    ids = _scheme_ids(scheme)
    shapes = _leaf_shapes(tensors, ids)
    options = options or default_options()
    key = (scheme_fingerprint(scheme), sparse, dtype, tuple(sorted(shapes.items())), tuple(slicing_bonds),
           repr(sorted((slicing_indices or {}).items(), key=repr)), repr(options))
    hit = _PLAN_CACHE.get(key)
    if hit is not None:
        return hit
    plan = ContractionPlan(scheme, shapes, sparse, slicing_bonds=slicing_bonds, slicing_indices=slicing_indices,
                           dtype=dtype, options=options)
    if len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
        _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
    _PLAN_CACHE[key] = plan
    return plan


_WORKSPACES = {}


def get_workspace(device, nbytes, stream=None):
    """One growing workspace (the arena of include/tnc_b200.h) per (device, stream): executions
    enqueued on different streams of one device never share an arena, executions on one stream are
    ordered by the stream.  `stream` is a torch.cuda.Stream, default the current one."""
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    if stream is None:
        stream = torch.cuda.current_stream(device)
    key = (device, int(stream.cuda_stream))
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        _WORKSPACES.pop(key, None)
        ws = None
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def release_workspaces():
    _WORKSPACES.clear()


def _device_of(tensors, ids):
    dev = tensors[ids[0]].device
    if dev.type != "cuda":
        raise RuntimeError(
            "artensor_b200 executes on CUDA devices only (no CPU fallback); got tensors on "
            f"{dev}. Move the leaf tensors to a B200 first.")
    return dev


def _run(tensors, scheme, sparse):
    if len(scheme) == 0:
        raise SchemeError("empty scheme")
    ids = _scheme_ids(scheme)
    dev = _device_of(tensors, ids)
    dt = tensors[ids[0]].dtype
    if dt not in _DTYPES:
        raise RuntimeError(f"artensor_b200: unsupported dtype {dt}; supported: {list(_DTYPES)}")
    plan = get_plan(scheme, tensors, sparse, options=mode_options(_DTYPES[dt]))
    with torch.cuda.device(dev):
        blob = plan.pack_leaves(tensors, device=dev)
        out = torch.zeros(plan.out_shape, dtype=torch.complex64, device=dev)
        ws = get_workspace(dev, plan.workspace_bytes)
        plan.execute(blob, out, 0, 1, ws, torch.cuda.current_stream(dev).cuda_stream)
    return plan, out


def tensor_contraction(tensors, scheme):
    """perform the tensor contraction (drop-in for artensor/contraction.py:62-76)"""
    plan, out = _run(tensors, scheme, sparse=False)
    tensors[plan.result_slot] = out
    return out


def tensor_contraction_sparse(tensors, contraction_scheme, scientific_notation=False):
    """drop-in for artensor/contraction.py:132-205.

    With scientific_notation=True the reference rescales after every step and returns
    (sum of log10 factors, rescaled tensor) (contraction.py:197-205).  Here the complex64 path
    needs no rescaling for range, so one factor is taken at the end: the returned pair satisfies
    the same contract, result == tensor * 10**factor with max|tensor| == 1."""
    plan, out = _run(tensors, contraction_scheme, sparse=True)
    for s in contraction_scheme:
        # consumed right operands are dropped like the reference does (contraction.py:174,188,191);
        # single-chunk batched steps keep tensors[j] there (:175-178: the gathered rows) -- here the
        # entry is left as it was, the gathered copy is never materialised
        if not (len(s) > 3 and len(s[2][0]) == 1 and len(s[2][1]) == 1):
            tensors[s[0][1]] = []
    tensors[plan.result_slot] = out
    if scientific_notation:
        nf = out.abs().max()
        factor = torch.log10(nf).to(out.dtype)
        out = out / nf
        tensors[plan.result_slot] = out
        return factor, out
    return out
