"""Multi-GPU plumbing of the path: rounds (independent leaf merges) are dealt to ranks by index, every rank
aligns its share on its own GPU, and the variable-length match lists go back to rank 0, which concatenates them in round
order for the serial graph merge.  The only collective is that gather (NCCL on GPUs; gloo in the CPU tests)."""
import struct

import torch
import torch.distributed as dist


def rounds_of_rank(n_rounds, rank, world):
    """Round r belongs to rank r % world (pair index modulo the number of GPUs)."""
    return list(range(rank, n_rounds, world))


def pack_rounds(payloads):
    """[(round_index, bytes)] -> one buffer: u64 count, then per round u64 index, u64 size, payload."""
    parts = [struct.pack("<Q", len(payloads))]
    for idx, blob in payloads:
        parts.append(struct.pack("<QQ", idx, len(blob)))
        parts.append(blob)
    return b"".join(parts)


def unpack_rounds(buf):
    (n,) = struct.unpack_from("<Q", buf, 0)
    off, out = 8, []
    for _ in range(n):
        idx, size = struct.unpack_from("<QQ", buf, off)
        off += 16
        out.append((idx, bytes(buf[off:off + size])))
        off += size
    return out


def gather_rounds(payloads, device, group=None):
    """Every rank contributes [(round_index, bytes)]; rank 0 gets all rounds sorted by index, the others get None."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    blob = pack_rounds(payloads)
    size = torch.tensor([len(blob)], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    mx = max(int(s.item()) for s in sizes)
    buf = torch.zeros(mx, dtype=torch.uint8, device=device)
    buf[:len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
    dst = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, dst, dst=0, group=group)
    if rank != 0:
        return None
    rounds = []
    for r in range(world):
        rounds.extend(unpack_rounds(dst[r][:int(sizes[r].item())].cpu().numpy().tobytes()))
    rounds.sort(key=lambda t: t[0])
    return rounds
