"""Multi-GPU host logic.  The path shards by independent units (images): every rank holds a full
weight replica and processes image i iff i % world == rank (static round-robin; the reference itself
is batch-1 single-GPU, densecap/DenseCapModel.lua:244, densecap/utils.lua:28).  The only exchange
is a gather of fixed-size padded result records, done with one all_gather
over NCCL (NVLink) -- or gloo in the CPU tests.  No data-path collective exists.
"""
from __future__ import annotations

import torch


def shard_indices(n_images: int, rank: int, world: int):
    """Static round-robin ownership: image i -> rank i % world."""
    return list(range(rank, n_images, world))


def record_numel(batch: int, max_proposals: int, seq_length: int) -> int:
    return batch + batch * max_proposals * (4 + 1 + seq_length)


def pack_result_record(num_final, final_boxes, final_scores, final_seq, out=None):
    """Packs one step's results into a flat fp32 record: [num_final (B) | boxes (B*P*4) | scores (B*P) | seq (B*P*T)].
    Token ids (<= V+1 ~ 1e4) and counts are exactly representable in fp32."""
    B = num_final.numel()
    rows = final_scores.numel()
    T = final_seq.numel() // max(rows, 1)
    n = B + rows * (5 + T)
    if out is None:
        out = torch.empty((n,), dtype=torch.float32, device=final_boxes.device)
    o = 0
    out[o:o + B] = num_final.to(torch.float32); o += B
    out[o:o + rows * 4] = final_boxes.reshape(-1); o += rows * 4
    out[o:o + rows] = final_scores.reshape(-1); o += rows
    out[o:o + rows * T] = final_seq.reshape(-1).to(torch.float32)
    return out


def unpack_result_record(rec, batch: int, max_proposals: int, seq_length: int):
    rows = batch * max_proposals
    o = 0
    num_final = rec[o:o + batch].to(torch.int32); o += batch
    boxes = rec[o:o + rows * 4].reshape(rows, 4); o += rows * 4
    scores = rec[o:o + rows]; o += rows
    seq = rec[o:o + rows * seq_length].reshape(rows, seq_length).to(torch.int64)
    return num_final, boxes, scores, seq


def gather_records(dist, record, out=None):
    """all_gather of equally sized records; returns a (world, numel) tensor on every rank."""
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world * record.numel(),), dtype=record.dtype, device=record.device)
    dist.all_gather_into_tensor(out, record)
    return out.view(world, record.numel())
