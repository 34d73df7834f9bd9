"""Multi-GPU host logic: utterances are independent (the reference never shares state across the
batch -- cc/kernels/ctc_ext_beam_search_decoder_kernels.cc:68-90 merely re-uses one decoder after
Reset()), so the batch is cut into contiguous blocks, one per GPU, and the sparse outputs are
concatenated on the host. There is NO collective on the decode path.

  * shard_bounds / merge_raw        pure host logic (numpy), used by both drivers below
  * decode_multi_device             one process, several GPUs (a host thread + stream per device)
  * decode_distributed              one process per GPU under torch.distributed (NCCL or gloo): each
                                    rank decodes its block; rank `dst` receives the merged result
"""
import threading

import numpy as np

from .decoder import CTCExtBeamSearchDecoder, ctc_ext_beam_search_decoder_raw


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
    return CTCExtBeamSearchDecoder(*groups, logp)


def decode_multi_device(inputs, sequence_length, beam_width, top_paths, merge_repeated=False,
                        blank_index=0, blank_label=-1, devices=None, decode_fn=None):
    """Single process, several GPUs: block r of the batch goes to devices[r]. Host arrays in, merged
    host (numpy) result out."""
    import torch
    x = _np(inputs)
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
            xs = np.ascontiguousarray(x[:, b0:b1, :])
            kw = {} if decode_fn else {"device": "cuda:%d" % devices[r]}
            results[r] = fn(xs, sl[b0:b1], beam_width, top_paths, merge_repeated, blank_index,
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


def decode_distributed(inputs, sequence_length, beam_width, top_paths, merge_repeated=False,
                       blank_index=0, blank_label=-1, dst=0, group=None, decode_fn=None):
    """One process per GPU: every rank holds the full (host) batch description, decodes its own
    contiguous block, and the raw outputs are gathered on rank `dst` (host-side gather of small
    sparse tensors; returns None elsewhere)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    x = _np(inputs)
    sl = _np(sequence_length).astype(np.int32)
    B = x.shape[1]
    bounds = shard_bounds(B, world)
    b0, b1 = bounds[rank]
    fn = decode_fn or ctc_ext_beam_search_decoder_raw
    err, mine = None, None
    try:
        raw = fn(np.ascontiguousarray(x[:, b0:b1, :]), sl[b0:b1], beam_width, top_paths,
                 merge_repeated, blank_index, blank_label)
        mine = tuple([_np(t) for t in g] for g in raw[:6]) + (_np(raw[6]),)
    except Exception as e:
        err = e
    gathered = [None] * world if rank == dst else None
    dist.gather_object((mine, err), gathered, dst=dst, group=group)
    if rank != dst:
        if err is not None:
            raise err
        return None
    for res, e in gathered:  # the reference aborts at the first failing utterance
        if e is not None:
            raise e
    return merge_raw([g[0] for g in gathered], bounds, B)
