"""Checkpoint path of the generator: the reference's ``<epoch>_net_G.pth`` files in and out, lean
device placement, and a packed-weight blob cache.

* ``save_network`` / ``load_network`` restate ``BaseModel.save_networks`` / ``load_networks`` (reference
  models/base_model.py:89-101, 114-131) for one network: unwrap ``DataParallel``, ``state_dict`` on the CPU, keys
  without a ``module.`` prefix, strict ``load_state_dict`` after dropping pre-0.4 InstanceNorm running statistics.
  Checkpoints written by the reference load here and vice versa (same 958 keys, same shapes).
* ``lean=True`` keeps the tensors the forward never reads (``_native.is_dead_key``: 208.6 M of 479.9 M elements at
  hidden_dim_ratio=4) in host memory: they still round-trip through ``state_dict``, but cost no HBM.
* ``save_packed`` / ``load_packed`` store / restore the plan's packed weight arena (fp16 operand panels, folded
  ActNorm affines) so that a deployment never has to hold the fp32 reference-layout tensors on the device at all.
"""
import hashlib

import torch

from . import _native

__all__ = ["save_network", "load_network", "to_device_lean", "save_packed", "load_packed", "packed_header"]


def _unwrap(net):
    return net.module if isinstance(net, torch.nn.DataParallel) else net


def save_network(net, path):
    """torch.save(net.module.cpu().state_dict()) without moving the live module off its device."""
    sd = {k: v.detach().cpu() for k, v in _unwrap(net).state_dict().items()}
    torch.save(sd, path)


def _patch_instance_norm(state_dict, net):
    """Drop ``running_mean`` / ``running_var`` entries of InstanceNorm layers that track no statistics (checkpoints
    written by PyTorch < 0.4; reference models/base_model.py:103-111)."""
    mods = dict(net.named_modules())
    for key in list(state_dict.keys()):
        owner, _, leaf = key.rpartition(".")
        m = mods.get(owner)
        if m is not None and type(m).__name__.startswith("InstanceNorm") and leaf in ("running_mean", "running_var"):
            if getattr(m, leaf, None) is None:
                state_dict.pop(key)


def load_network(net, path, device=None, lean=False):
    """Load ``<epoch>_net_G.pth`` into ``net`` (strict).  ``device``: where the live tensors go afterwards (default:
    where they are); ``lean``: leave the dead tensors on the host."""
    target = _unwrap(net)
    state_dict = torch.load(path, map_location="cpu")
    if any(k.startswith("module.") for k in state_dict):
        state_dict = {k[len("module."):] if k.startswith("module.") else k: v for k, v in state_dict.items()}
    _patch_instance_norm(state_dict, target)
    was = next((p.device for n, p in target.named_parameters() if not _native.is_dead_key(n)), torch.device("cpu"))
    if was.type != "cpu":
        target.to("cpu")        # load on the host, then place (lean placement must not see stale device copies)
    target.load_state_dict(state_dict, strict=True)
    dev = torch.device(device) if device is not None else was
    if dev.type != "cpu":
        to_device_lean(target, dev) if lean else target.to(dev)
    return net


def to_device_lean(net, device):
    """Move every tensor the forward reads to ``device``; dead tensors stay where they are (host)."""
    target = _unwrap(net)
    device = torch.device(device)
    for mod_name, mod in target.named_modules():
        for store in (mod._parameters, mod._buffers):
            for leaf, t in store.items():
                if t is None:
                    continue
                full = (mod_name + "." if mod_name else "") + leaf
                if not _native.is_dead_key(full):
                    t.data = t.data.to(device)
    target.mark_weights_changed()
    return net


def packed_header(net, height, width):
    """What a packed blob is valid for: library ABI, plan configuration, and a digest of the state_dict it was packed
    from (keys, shapes and bytes of the live tensors)."""
    target = _unwrap(net)
    opt = target.opt
    h = hashlib.sha256()
    for k, v in target.state_dict().items():
        if _native.is_dead_key(k) or not v.is_floating_point():
            continue
        h.update(k.encode())
        h.update(str(tuple(v.shape)).encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return {"abi": _native.lib().cfen_abi_version(), "height": int(height), "width": int(width),
            "n_feats": int(opt.n_feats), "hidden_dim_ratio": int(opt.hidden_dim_ratio), "num_heads": int(opt.num_heads),
            "patch_dim": int(opt.patch_dim), "patch_size": int(opt.patch_size), "loadSize": int(opt.loadSize),
            "weights_sha256": h.hexdigest()}


def save_packed(net, path, height, width):
    """Pack (if needed) on the device the live tensors are on and write header + blob."""
    target = _unwrap(net)
    dev = next(p.device for n, p in target.named_parameters() if not _native.is_dead_key(n))
    with torch.cuda.device(dev):
        plan = target._plan_for(dev, height, width, 1)
        if plan.needs_calibration:
            raise RuntimeError("save_packed needs initialised ActNorm layers (run one forward or load a checkpoint)")
        blob = plan.export_packed(torch.cuda.current_stream(dev).cuda_stream)
    torch.save({"header": packed_header(target, height, width), "blob": blob}, path)


def load_packed(net, path, device, height, width, max_batch=1, verify=True):
    """Fill a fresh plan on ``device`` from a packed blob.  The module's own tensors may stay on the host: the forward
    reads only the plan.  ``verify`` recomputes the header from ``net`` (its state_dict must be the one the blob was
    packed from) and refuses a mismatch."""
    target = _unwrap(net)
    device = torch.device(device)
    rec = torch.load(path, map_location="cpu")
    want = packed_header(target, height, width) if verify else None
    if verify and rec["header"] != want:
        diff = sorted(k for k in want if rec["header"].get(k) != want[k])
        raise RuntimeError("packed blob does not match this network / library (differs in: %s)" % ", ".join(diff))
    opt = target.opt
    with torch.cuda.device(device):
        plan = _native.Plan(device=device.index, height=height, width=width, n_feats=opt.n_feats,
                            hidden_dim_ratio=opt.hidden_dim_ratio, num_heads=opt.num_heads, patch_dim=opt.patch_dim,
                            patch_size=int(opt.patch_size), load_size=int(opt.loadSize),
                            max_batch=_native.chunk_images(height, max_batch))
        plan.import_packed(rec["blob"], torch.cuda.current_stream(device).cuda_stream)
    plan.weights_version = (target._weights_version, target._tensors_version())
    target._plans[(device.index, height, width)] = plan
    return plan
