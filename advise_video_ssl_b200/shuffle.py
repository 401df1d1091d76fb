"""Shuffle-BN bookkeeping (A6: models/contrastive.py:174-230) derived on the HOST.

The reference draws `randperm(W*B)` on rank 0's CPU generator, moves it to the GPU, broadcasts
it and all_gathers the whole key clip (C1: 308 MB in, 2.47 GB out per rank at cfg2) to keep B
rows of it.  Here only the rows a rank keeps cross the fabric (an all-to-all with split sizes
known on every host), and nothing on the way needs a device->host copy:

  * the permutation travels between the HOSTS (4 KB over a gloo side group, or the process
    group itself when it can move CPU tensors), so every rank can derive its send/receive
    lists without waiting for its CUDA stream;
  * the index tensors the exchange needs are uploaded from pinned memory.

`ShufflePlan` holds everything both directions need for one draw.
"""
import numpy as np
import torch
import torch.distributed as dist

_side_groups = {}


def _cpu_capable(group=None):
    try:
        return "gloo" in str(dist.get_backend(group)).lower() or "mpi" in str(dist.get_backend(group)).lower()
    except Exception:  # noqa: BLE001
        return False


def _host_group():
    """A gloo group over WORLD for small host-side broadcasts, created once (collective: the first
    shuffled step reaches this on every rank at the same point).  None when the default group can
    move CPU tensors itself."""
    if _cpu_capable():
        return None
    g = _side_groups.get("world")
    if g is None:
        g = dist.new_group(backend="gloo")
        _side_groups["world"] = g
    return g


def broadcast_from_rank0(values_cpu, device):
    """In-place broadcast of a small CPU tensor from global rank 0 (C2).  Host-to-host, so no CUDA
    stream is drained; if no CPU-capable group can be built the broadcast runs on the device and is
    read back (one synchronisation, the reference's own behaviour at :199-201 + :265)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return values_cpu
    state = _side_groups.get("mode")
    if state != "device":
        try:
            dist.broadcast(values_cpu, src=0, group=_host_group())
            _side_groups["mode"] = "host"
            return values_cpu
        except Exception:  # noqa: BLE001 - no gloo in this build: fall back for good, on every rank alike
            if state == "host":
                raise
            _side_groups["mode"] = "device"
    on_dev = values_cpu.to(device)
    dist.broadcast(on_dev, src=0)
    values_cpu.copy_(on_dev.cpu())
    return values_cpu


class ShufflePlan:
    """One draw of the shuffle permutation and what follows from it.

    perm      int64 [W*B] on the host: row r of rank d's shuffled batch is row perm[d*B + r] of the
              rank-major concatenation of all local batches (reference :204-207).
    restore   int64 [W, B] on the device = argsort(perm).view(W, B): `_batch_unshuffle`'s index
              (:209-212); row i of rank d's ORIGINAL batch sits at gathered[restore[d, i]].  Row
              `rank` of it is also where this rank's local rows GO in the shuffled batch (the scatter's
              destination positions).
    For the NCCL all-to-all: `send_rows` (device int64: local rows in the order they are sent, grouped
    by destination), `send_counts` / `recv_counts` (host lists) and `place` (device int64: where each
    row of the shuffled batch sits in the receive buffer).
    Every index array travels to the device in ONE copy from one pinned buffer, which the plan keeps
    alive (the copy is asynchronous and may be a captured graph node).
    """

    def __init__(self, perm, world, rank, bsz, device, need_alltoall=True):
        perm = np.asarray(perm, dtype=np.int64).reshape(world, bsz)
        self.world, self.rank, self.bsz, self.device = world, rank, bsz, device
        self.perm = perm
        restore_host = np.argsort(perm.reshape(-1), kind="stable")
        self.dest_host = np.ascontiguousarray(restore_host.reshape(world, bsz)[rank])  # where this rank's rows go
        parts = [perm[rank], restore_host]
        if world > 1 and need_alltoall:
            holder = perm // bsz  # rank that owns each wanted row
            # np.nonzero walks row-major: destinations in ascending order, each in its own take order
            dst, pos = np.nonzero(holder == rank)
            self.send_counts = np.bincount(dst, minlength=world).tolist()
            self.recv_counts = np.bincount(holder[rank], minlength=world).tolist()
            # the receive buffer is grouped by source rank; inside a group rows keep this rank's take order
            arrival = np.argsort(holder[rank], kind="stable")
            place = np.empty(bsz, dtype=np.int64)
            place[arrival] = np.arange(bsz)
            parts += [perm[dst, pos] % bsz, place]
        self._n_parts = len(parts)
        self._host = torch.from_numpy(np.ascontiguousarray(np.concatenate(parts)))
        if device.type == "cuda":
            self._host = self._host.pin_memory()
        self._dev = None

    def _packed(self):
        """The index arrays on the device: ONE asynchronous copy from the pinned buffer, issued on first use (the
        NVLink scatter carries its positions in the kernel parameters and runs before it)."""
        if self._dev is None:
            self._dev = self._host.to(self.device, non_blocking=True) if self.device.type == "cuda" else self._host
        return self._dev

    @property
    def take(self):
        return self._packed()[:self.bsz]

    @property
    def restore(self):
        return self._packed()[self.bsz:(self.world + 1) * self.bsz].view(self.world, self.bsz)

    @property
    def send_rows(self):
        assert self._n_parts == 4
        return self._packed()[(self.world + 1) * self.bsz:(self.world + 2) * self.bsz]

    @property
    def place(self):
        assert self._n_parts == 4
        return self._packed()[(self.world + 2) * self.bsz:]

    def shuffled(self, x, group=None, scatter=None, status=None):
        """This rank's shuffled batch: cat_all_gather(x)[perm[rank]] bit for bit, moving B rows.
        `scatter`: an ops.PeerScatter for x's row size -> NVLink peer stores instead of NCCL."""
        if self.world == 1:
            return x.index_select(0, self.take)
        if scatter is not None:
            small = self.bsz <= 256  # positions ride in the kernel parameters: no dependence on the index upload
            return scatter.exchange(x.contiguous(), None if small else self.restore[self.rank], status=status,
                                    dest_pos_host=self.dest_host)
        outgoing = x.index_select(0, self.send_rows)
        incoming = torch.empty_like(x)
        dist.all_to_all_single(incoming, outgoing, output_split_sizes=self.recv_counts,
                               input_split_sizes=self.send_counts, group=group)
        return incoming.index_select(0, self.place)
