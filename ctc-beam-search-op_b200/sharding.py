"""Multi-GPU host logic: utterances are independent (the reference never shares state across the
batch -- cc/kernels/ctc_ext_beam_search_decoder_kernels.cc:68-90 merely re-uses one decoder after
Reset()), so the batch is cut into contiguous blocks, one per GPU, each block is decoded IN PLACE from
a view `inputs[:, b0:b1, :]` of the time-major tensor (no repack: the kernels take a time stride, a
host view is copied with a pitched copy), and the sparse outputs are concatenated. There is NO
collective on the decode path; the only exchange is the gather of the (small) sparse outputs.

  * shard_bounds / merge_raw        pure host logic, used by both drivers below
  * decode_multi_device             one process, several GPUs (a host thread + stream per device)
  * decode_distributed              one process per GPU under torch.distributed: each rank decodes its
                                    block; rank `dst` receives the merged result. Under NCCL the packed
                                    outputs travel GPU -> GPU (NVLink) and are merged on the device;
                                    under gloo (CPU tests) they are gathered as host objects.
"""
import threading

import numpy as np

from . import decoder as _dec
from .decoder import CTCExtBeamSearchDecoder, ctc_ext_beam_search_decoder_raw


def bind_host_to_device(device=None):
    """One process per GPU: run this process (and, by first touch, the pinned buffers it allocates from
    now on) on the CPU cores next to `device` -- /sys/bus/pci/devices/<bdf>/local_cpulist. On a
    two-socket box the host->device feed of the decode otherwise crosses the socket link for half of the
    ranks. Returns the cpu list, or None when the topology cannot be read (nothing is changed then)."""
    import os

    import torch
    try:
        idx = torch.device("cuda", torch.cuda.current_device() if device is None else device).index \
            if not isinstance(device, int) else device
        pr = torch.cuda.get_device_properties(idx)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as fh:
            text = fh.read().strip()
        cpus = set()
        for part in text.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)  # stay inside what the container allows
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def shard_bounds(batch, parts):
    """Contiguous, balanced blocks: [(b0, b1)] * parts (empty blocks allowed when batch < parts)."""
    base, extra = divmod(int(batch), int(parts))
    out, b0 = [], 0
    for r in range(parts):
        b1 = b0 + base + (1 if r < extra else 0)
        out.append((b0, b1))
        b0 = b1
    return out


def _np(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def merge_raw(shards, bounds, batch):
    """Concatenate per-shard raw outputs (7 groups) in shard order. Row indices are shifted by the
    shard's first utterance; dense_shape = [batch, max over shards] -- exactly what
    StoreAllDecodedSequences (kernels.cc:163-257) produces for the whole batch."""
    P = len(shards[0][0])
    groups = [[], [], [], [], [], []]
    for p in range(P):
        for base in (0, 3):
            idx, val, mx = [], [], 0
            for sh, (b0, _) in zip(shards, bounds):
                i = _np(sh[base][p]).astype(np.int64).reshape(-1, 2).copy()
                i[:, 0] += b0
                idx.append(i)
                val.append(_np(sh[base + 1][p]).astype(np.int64).reshape(-1))
                mx = max(mx, int(_np(sh[base + 2][p])[1]))
            groups[base].append(np.concatenate(idx, axis=0))
            groups[base + 1].append(np.concatenate(val, axis=0))
            groups[base + 2].append(np.asarray([batch, mx], np.int64))
    logp = np.concatenate([_np(sh[6]).reshape(-1, P) for sh in shards], axis=0)
    res = CTCExtBeamSearchDecoder(*groups, logp)
    res.flags = 0
    for sh in shards:
        res.flags |= int(getattr(sh, "flags", 0))
    return res


def _merge_device(shards, counts, maxima, bounds, batch, f64):
    """merge_raw for torch tensors that live on one device (no host round trip): the merged outputs are
    written straight into ONE packed int64 buffer (the layout of decoder._carve), so a host caller gets
    them with a single copy into page-locked memory."""
    import torch
    P = len(shards[0][0])
    dev = shards[0][6].device
    n_dec = [sum(c[0][p] for c in counts) for p in range(P)]
    n_ali = [sum(c[1][p] for c in counts) for p in range(P)]
    n = _dec._pack_elems(batch, P, (n_dec, n_ali), f64)
    buf = torch.empty((n,), dtype=torch.int64, device=dev)
    groups, logp = _dec._carve(buf, batch, P, (n_dec, n_ali), f64)
    shapes = torch.empty((2 * P, 2), dtype=torch.int64)
    for p in range(P):
        for k, base in enumerate((0, 3)):
            idx, val = [], []
            for sh, (b0, _) in zip(shards, bounds):
                i = sh[base][p]
                if b0:
                    i[:, 0] += b0  # in place: the gathered buffer is ours
                idx.append(i)
                val.append(sh[base + 1][p])
            if groups[base][p].numel():
                torch.cat(idx, dim=0, out=groups[base][p])
                torch.cat(val, dim=0, out=groups[base + 1][p])
            shapes[2 * p + k, 0] = batch
            shapes[2 * p + k, 1] = max(m[k][p] for m in maxima)  # the longest sequence over all shards (host values)
    shapes = shapes.to(dev, non_blocking=True)
    for p in range(P):
        groups[2][p].copy_(shapes[2 * p])
        groups[5][p].copy_(shapes[2 * p + 1])
    if logp.numel():
        torch.cat([sh[6].reshape(-1, P) for sh in shards], dim=0, out=logp)
    res = CTCExtBeamSearchDecoder(*groups, logp)
    res.packed = buf
    return res, (n_dec, n_ali)


def _view(x, b0, b1):
    """Block [b0, b1) of the batch axis as a VIEW (numpy or torch): no copy."""
    return x[:, b0:b1, :]


def decode_multi_device(inputs, sequence_length, beam_width, top_paths, merge_repeated=False,
                        blank_index=0, blank_label=-1, devices=None, decode_fn=None):
    """Single process, several GPUs: block r of the batch goes to devices[r]. Host arrays in, merged
    host (numpy) result out."""
    import torch
    x = inputs if isinstance(inputs, (np.ndarray, torch.Tensor)) else np.asarray(inputs)
    sl = _np(sequence_length).astype(np.int32)
    B = x.shape[1]
    if devices is None:
        devices = list(range(torch.cuda.device_count()))
    bounds = shard_bounds(B, len(devices))
    fn = decode_fn or ctc_ext_beam_search_decoder_raw
    results, errors = [None] * len(devices), [None] * len(devices)

    def work(r):
        b0, b1 = bounds[r]
        try:
            kw = {} if decode_fn else {"device": "cuda:%d" % devices[r], "batch_offset": b0}
            results[r] = fn(_view(x, b0, b1), sl[b0:b1], beam_width, top_paths, merge_repeated, blank_index,
                            blank_label, **kw)
        except Exception as e:  # re-raised below, first failing shard first (reference order)
            errors[r] = e

    threads = [threading.Thread(target=work, args=(r,)) for r in range(len(devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None:
            raise e
    return merge_raw(results, bounds, B)


def _gather_packed_nccl(raw, err, bounds, B, P, f64, rank, world, dst, group, to_host, max_time):
    """Gather of the packed outputs GPU -> GPU. Every decode leaves its 6*P + 1 outputs as views of ONE
    int64 buffer (`raw.packed`); the buffers (padded to the longest) are gathered on `dst` with one
    NCCL call, carved up there and merged on the device."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device())
    # header: [error code, error batch index, flags, n_dec[P], n_ali[P], packed length, max_dec[P], max_ali[P]]
    hdr = torch.zeros(4 + 4 * P, dtype=torch.int64)
    if err is not None:
        hdr[0] = int(getattr(err, "code", -1)) or -1
        hdr[1] = int(getattr(err, "batch_index", -1))
    else:
        hdr[2] = int(raw.flags)
        for p in range(P):
            hdr[3 + p] = raw[0][p].shape[0]
            hdr[3 + P + p] = raw[3][p].shape[0]
        hdr[3 + 2 * P] = raw.packed.numel()
        for p in range(P):
            hdr[4 + 2 * P + p] = raw.max_lengths[0][p]
            hdr[4 + 3 * P + p] = raw.max_lengths[1][p]
    hdr = hdr.to(dev)
    all_hdr = torch.empty((world, hdr.numel()), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_hdr, hdr, group=group)
    all_hdr = all_hdr.cpu()
    failed = [r for r in range(world) if int(all_hdr[r, 0]) != 0]
    if failed:  # every rank learns of the failure; the reference aborts at the first failing utterance
        if err is not None and (rank != dst or failed[0] == rank):
            raise err
        r0 = failed[0]
        if rank == dst:
            code, b = int(all_hdr[r0, 0]), int(all_hdr[r0, 1])
            lib = _dec._lib.load()
            if code == 5:  # kernels.cc:134-138, with the index in the whole batch
                raise _dec.FailedPreconditionError(5, "sequence_length(%d) <= %d" % (b, max_time))
            if code in _dec._ERR_CLASS:
                raise _dec._ERR_CLASS[code](code, lib.ctcx_strerror(code).decode())
            raise RuntimeError("ctcx: rank %d failed (code %d)" % (r0, code))
        return None
    n_max = int(all_hdr[:, 3 + 2 * P].max())
    mine = raw.packed
    if mine.numel() < n_max:
        padded = torch.empty(n_max, dtype=torch.int64, device=dev)
        padded[:mine.numel()] = mine
        mine = padded
    if rank == dst:
        recv = [torch.empty(n_max, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.gather(mine, recv, dst=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
    else:
        dist.gather(mine, None, dst=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
        return None
    shards, all_counts, maxima = [], [], []
    rows = all_hdr.tolist()
    for r in range(world):
        b0, b1 = bounds[r]
        counts = (rows[r][3:3 + P], rows[r][3 + P:3 + 2 * P])
        groups, logp = _dec._carve(recv[r], b1 - b0, P, counts, f64)
        shards.append(CTCExtBeamSearchDecoder(*groups, logp))
        all_counts.append(counts)
        maxima.append((rows[r][4 + 2 * P:4 + 3 * P], rows[r][4 + 3 * P:4 + 4 * P]))
    out, merged_counts = _merge_device(shards, all_counts, maxima, bounds, B, f64)
    flags = 0
    for r in range(world):
        flags |= int(all_hdr[r, 2])
    if to_host:  # one copy of the packed result into page-locked memory, carved up there
        hbuf = torch.empty(out.packed.shape, dtype=torch.int64, pin_memory=True)
        hbuf.copy_(out.packed, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        groups, logp = _dec._carve(hbuf, B, P, merged_counts, f64)
        out = CTCExtBeamSearchDecoder(*groups, logp)
        out.packed = hbuf
    out.flags = flags
    return out


def decode_distributed(inputs, sequence_length, beam_width, top_paths, merge_repeated=False,
                       blank_index=0, blank_label=-1, dst=0, group=None, decode_fn=None, global_batch=None):
    """One process per GPU: each rank decodes its contiguous block of the batch and the raw outputs are
    gathered on rank `dst` (returns None elsewhere). Outputs live where the inputs live (device tensors
    in -> device tensors on `dst`).

    global_batch=None   every rank passes the WHOLE batch ([T, B, C], [B]); only the view
                        `inputs[:, b0:b1, :]` of its own block is read -- in place, no repack.
    global_batch=B      every rank passes only ITS block ([T, b1-b0, C], [b1-b0]) of a batch of B
                        utterances cut by `shard_bounds(B, world_size)` -- the usual layout of a
                        multi-process job."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    x = inputs if isinstance(inputs, (np.ndarray, torch.Tensor)) else np.asarray(inputs)
    sl = sequence_length if isinstance(sequence_length, torch.Tensor) else np.asarray(sequence_length, np.int32)
    local = global_batch is not None
    B = int(global_batch) if local else (int(x.shape[1]) if x.ndim == 3 else 0)
    bounds = shard_bounds(B, world)
    b0, b1 = bounds[rank]
    if local and x.ndim == 3 and int(x.shape[1]) != b1 - b0:
        raise ValueError("rank %d holds %d utterances, its block of a batch of %d over %d ranks has %d"
                         % (rank, int(x.shape[1]), B, world, b1 - b0))
    device_gather = decode_fn is None and dist.get_backend(group) == "nccl" and x.ndim == 3
    err, raw = None, None
    try:
        xb, slb = (x, sl) if local else (_view(x, b0, b1), sl[b0:b1])
        if decode_fn is not None:
            raw = decode_fn(np.ascontiguousarray(_np(xb)), _np(slb).astype(np.int32), beam_width,
                            top_paths, merge_repeated, blank_index, blank_label)
        else:
            raw = ctc_ext_beam_search_decoder_raw(xb, slb, beam_width, top_paths,
                                                  merge_repeated, blank_index, blank_label, batch_offset=b0,
                                                  outputs="device" if device_gather else "auto")
    except Exception as e:
        err = e
    if device_gather:
        f64 = (x.dtype == torch.float64) if isinstance(x, torch.Tensor) else (x.dtype == np.float64)
        to_host = not (isinstance(x, torch.Tensor) and x.is_cuda)
        out = _gather_packed_nccl(raw, err, bounds, B, int(top_paths), f64, rank, world, dst, group, to_host,
                                  int(x.shape[0]))
        if out is not None and to_host and not isinstance(x, torch.Tensor):  # numpy in -> numpy out
            flags = out.flags
            out = CTCExtBeamSearchDecoder(*[[t.numpy() for t in g] for g in out[:6]], out[6].numpy())
            out.flags = flags
        return out
    mine = None
    if raw is not None:
        mine = tuple([_np(t) for t in g] for g in raw[:6]) + (_np(raw[6]), int(getattr(raw, "flags", 0)))
    gathered = [None] * world if rank == dst else None
    dist.gather_object((mine, err), gathered, dst=dst, group=group)
    if rank != dst:
        if err is not None:
            raise err
        return None
    for res, e in gathered:  # the reference aborts at the first failing utterance
        if e is not None:
            raise e
    shards = []
    for g, _ in gathered:
        sh = CTCExtBeamSearchDecoder(*g[:7])
        sh.flags = g[7]
        shards.append(sh)
    return merge_raw(shards, bounds, B)
