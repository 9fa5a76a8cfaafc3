"""One-process-per-GPU driver for a frame: trajectory sharding + row-stripe exchange.

The reference's only parallelism is render_parallel (lib.rs:1051-1082): N workers each own a
private Runtime, take jobs off a counter, and the caller merges the N Runtimes and colourises.
Here a worker is a GPU (a rank), its private Runtime lives in its HBM, and the merge is done
stripe-wise by the GPUs themselves (include/sar.h, "One frame of render_parallel over N ranks"):

    reset      (first waits until every peer has finished reading this rank's previous frame)
    render     rank r renders jobs [r*J, (r+1)*J), order keys global       (no communication)
    export     counts in pixel order for the peers; RENDER_DONE to every rank
    merge      waits for everyone's RENDER_DONE; rank r reduces ROW STRIPE r over all ranks by
               loading the peers' counts and Δp records directly over NVLink (CUDA IPC mappings):
               counts add, the record with the greatest (z, earlier job) wins — Runtime::merge
               (lib.rs:708-738) made order-independent, so the frame is bit-identical for 1, 2, 4,
               8 ranks; the stripe's share of Runtime.max and of the Depth min/max goes to every
               rank; MAX_READY + MERGE_DONE
    colourise  waits for every stripe's maximum (the log base of lib.rs:860 and the Depth range of
               lib.rs:877-882 are global) and for the owner's IMAGE_FREE; rank r colourises stripe
               r straight into rank 0's image over NVLink; IMAGE_DONE to rank 0
    rank 0     waits for IMAGE_DONE from all, uses the image, raises IMAGE_FREE

All of that synchronisation is device-side — flags in the ranks' exported allocations, written by
remote stores and polled locally, as prologues / epilogues of those five kernels — the host only
enqueues.  torch.distributed is plumbing for set-up: the rendezvous and the all-gather of the
64-byte IPC handles (plus max-over-ranks of timings in bench.py).  Pure functions at the top
(job_slice, stripe_rows, iterations_per_job) are the host logic; they are covered by CPU tests
(gloo, world_size 2).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Tuple


# ---- pure host logic --------------------------------------------------------------------
def iterations_per_job(total_iterations: int, num_threads: int, jobs_per_thread: int) -> int:
    """lib.rs:1058: `iterations / num_threads / jobs_per_thread`, integer division twice."""
    return total_iterations // num_threads // jobs_per_thread


def job_slice(rank: int, world: int, lanes_per_rank: int, jobs_per_thread: int) -> Tuple[int, int]:
    """(first_job, n_jobs) of `rank`: contiguous slices of the num_threads*jobs_per_thread jobs
    (lib.rs:1062), num_threads = world*lanes_per_rank."""
    n = lanes_per_rank * jobs_per_thread
    return rank * n, n


def stripe_rows(rank: int, world: int, height: int) -> Tuple[int, int]:
    """(row0, rows) of the contiguous row stripe `rank` reduces and colourises.  Row-major
    storage (idx = y*w + x, as image::ImageBuffer) makes a stripe one contiguous range."""
    row0 = rank * height // world
    row1 = (rank + 1) * height // world
    return row0, row1 - row0


# ---- torch.distributed plumbing ---------------------------------------------------------
def init_process_group(world: int, rank: int, local_rank: int, backend: Optional[str] = None):
    import torch
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if not dist.is_initialized():
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device(f"cuda:{local_rank}")
        dist.init_process_group(backend=backend, world_size=world, rank=rank, **kw)
    return dist.group.WORLD


def barrier(group) -> None:
    import torch.distributed as dist

    dist.barrier(group=group)


def shutdown(group) -> None:
    import torch.distributed as dist

    if dist.is_initialized():
        dist.destroy_process_group()


def _device_of(group, local_rank: int):
    import torch
    import torch.distributed as dist

    return torch.device(f"cuda:{local_rank}") if dist.get_backend(group) == "nccl" else torch.device("cpu")


def max_over_ranks(value: float, group, local_rank: int) -> float:
    if group is None:
        return value
    import torch
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.float64, device=_device_of(group, local_rank))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def allreduce_max_u32(value: int, group, local_rank: int) -> int:
    """Runtime.max over all stripes (the log base of lib.rs:860 is global)."""
    if group is None:
        return value
    import torch
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.int64, device=_device_of(group, local_rank))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def allgather_bytes(payload: bytes, group, local_rank: int) -> List[bytes]:
    """All-gather of a fixed-size byte string (the 64-byte CUDA IPC handle of each rank)."""
    import torch
    import torch.distributed as dist

    dev = _device_of(group, local_rank)
    mine = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    outs = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
    dist.all_gather(outs, mine, group=group)
    return [bytes(o.cpu().tolist()) for o in outs]


# ---- a frame on this rank's GPU -----------------------------------------------------------
class Frame:
    """Owns this rank's Runtime and peer mappings; step_device() renders one whole frame."""

    def __init__(self, cfg, device: int, world: int, rank: int, group, lanes: int, jobs_per_thread: int,
                 iterations_per_gpu: int, seed: int):
        from . import _native as N

        self.N, self.L = N, N.lib()
        self.world, self.rank, self.group, self.device, self.seed = world, rank, group, device, seed
        if not lanes:
            t = C.c_uint32()
            N.check(N.lib().sar_default_threads(device, C.byref(t)))
            lanes = int(t.value)
        self.lanes = lanes
        self.jpt = jobs_per_thread
        self.total_iterations = iterations_per_gpu * world
        self.iterations_per_job = iterations_per_job(self.total_iterations, self.lanes * world, jobs_per_thread)
        self.first_job, self.n_jobs = job_slice(rank, world, self.lanes, jobs_per_thread)
        self.cfg_total = cfg.to_pod()
        self.cfg_total.iterations = self.total_iterations
        self.pod = cfg.to_pod()
        self.pod.iterations = self.iterations_per_job
        self.h, self.w = self.pod.height, self.pod.width
        self.row0, self.rows = stripe_rows(rank, world, self.h)
        self.rt = C.c_void_p()
        N.check(self.L.sar_runtime_new(self.w, self.h, device, C.byref(self.rt)))
        self.peers: List[Optional[C.c_void_p]] = [None] * world
        if world > 1:
            buf = (C.c_uint8 * N.SAR_IPC_HANDLE_BYTES)()
            N.check(self.L.sar_runtime_ipc_export(self.rt, buf))
            handles = allgather_bytes(bytes(buf), group, device)
            for r, hb in enumerate(handles):
                if r == rank:
                    continue
                p = C.c_void_p()
                hbuf = (C.c_uint8 * N.SAR_IPC_HANDLE_BYTES).from_buffer_copy(hb)
                N.check(self.L.sar_peer_open(hbuf, self.w, self.h, device, C.byref(p)))
                self.peers[r] = p
            self.peer_arr = (C.c_void_p * world)(*[(p.value if p is not None else None) for p in self.peers])
            barrier(group)                       # every rank has mapped every peer before the first remote store
        else:
            self.peer_arr = None
        self.epoch = 0
        self._renderer = None

    # -- bookkeeping
    def recorded_iterations_local(self) -> int:
        return self.iterations_per_job * self.n_jobs

    def recorded_iterations_total(self) -> int:
        return self.iterations_per_job * self.n_jobs * self.world

    # -- pieces
    def reset_async(self, sp) -> None:
        if self.world == 1:
            self.N.check(self.L.sar_runtime_reset_async(self.rt, sp))
        else:   # first waits (on the device) until every peer has read this rank's previous frame
            self.N.check(self.L.sar_frame_reset_async(self.rt, self.world, self.epoch, sp))

    def render_async(self, sp, d_init=None) -> None:
        # order keys are global over the ranks: rank r's jobs follow rank r-1's (include/sar.h)
        self.N.check(self.L.sar_runtime_set_job_base(self.rt, self.first_job))
        if d_init is None:
            self.N.check(self.L.sar_render_seeded_async(C.byref(self.pod), self.rt, self.seed, self.first_job,
                                                        self.n_jobs, self.lanes, sp))
        else:
            self.N.check(self.L.sar_render_device_async(C.byref(self.pod), self.rt, d_init, self.first_job,
                                                        self.n_jobs, self.lanes, sp))

    def finish_async(self, sp) -> None:
        """merge (N>1) → max → colourise; leaves the image in rank 0's HBM."""
        N, L = self.N, self.L
        if self.world == 1:
            N.check(L.sar_runtime_max_async(self.rt, 0, 0, sp))
            N.check(L.sar_colorize_rows_async(C.byref(self.pod), self.rt, 0, 0, None, sp))
            return
        e, n, r = self.epoch, self.world, self.rank
        N.check(L.sar_frame_export_async(self.rt, self.peer_arr, n, r, e, sp))
        N.check(L.sar_frame_merge_async(self.rt, self.peer_arr, n, r, self.row0, self.rows, e, sp))
        N.check(L.sar_frame_colorize_async(C.byref(self.pod), self.rt, self.peer_arr, n, r, 0, self.row0, self.rows, e, sp))
        if r == 0:
            N.check(L.sar_frame_image_wait_async(self.rt, n, e, sp))                     # rank 0's image is complete

    def release_image(self, sp) -> None:
        """Rank 0 is done with the frame's image (copied out, or not needed): peers may overwrite it."""
        if self.world > 1 and self.rank == 0:
            self.N.check(self.L.sar_frame_image_release_async(self.rt, self.peer_arr, self.world, 0, self.epoch, sp))

    def begin_frame(self, sp) -> None:
        self.epoch += 1

    def step_device(self, sp) -> None:
        self.begin_frame(sp)
        self.reset_async(sp)
        self.render_async(sp)
        self.finish_async(sp)
        self.release_image(sp)

    def check_sync(self) -> None:
        """Raise if a cross-GPU wait of an earlier frame timed out (its results are then void)."""
        if self.world == 1:
            return
        err = C.c_uint32()
        self.N.check(self.L.sar_runtime_sync_error(self.rt, C.byref(err), 0))
        if err.value:
            raise RuntimeError(f"cross-GPU wait timed out (kind {err.value - 1}) on rank {self.rank}")

    # -- end to end with host buffers
    def make_e2e(self):
        return _E2E(self)

    def close(self) -> None:
        for p in self.peers:
            if p is not None:
                self.L.sar_peer_close(p)
        self.peers = []
        if self._renderer is not None:
            self.L.sar_renderer_shutdown(self._renderer)
            self._renderer = None
        if self.rt:
            self.L.sar_runtime_free(self.rt)
            self.rt = None


class _E2E:
    """The frame through host buffers: start points in from pinned memory, RGBA16 image out."""

    def __init__(self, frame: Frame):
        import torch

        from . import api

        self.f = f = frame
        N, L = f.N, f.L
        pts = api.seed_points(f.seed, f.first_job, f.n_jobs)
        self.h_pts = torch.from_numpy(pts).pin_memory()
        self.h2d_bytes = self.h_pts.numel() * 8
        self.h_img = torch.empty((f.h, f.w, 4), dtype=torch.uint16).pin_memory() if f.rank == 0 else None
        self.d2h_bytes = f.h * f.w * 8
        if f.world == 1:
            dev = (C.c_int * 1)(f.device)
            f._renderer = C.c_void_p()
            N.check(L.sar_renderer_new(dev, 1, f.lanes, C.byref(f._renderer)))
        else:
            self.d_pts = torch.empty_like(self.h_pts, device=f"cuda:{f.device}")
            self.stream = torch.cuda.Stream(device=f.device)

    def step(self) -> None:
        import torch

        f = self.f
        N, L = f.N, f.L
        if f.world == 1:
            N.check(L.sar_render_parallel(f._renderer, C.byref(f.cfg_total), f.jpt, f.seed,
                                          C.cast(self.h_pts.data_ptr(), N._f64p), C.cast(self.h_img.data_ptr(), N._u16p)))
            return
        sp = C.c_void_p(self.stream.cuda_stream)
        with torch.cuda.stream(self.stream):
            self.d_pts.copy_(self.h_pts, non_blocking=True)
        f.begin_frame(sp)
        f.reset_async(sp)
        f.render_async(sp, C.c_void_p(self.d_pts.data_ptr()))
        f.finish_async(sp)
        if f.rank == 0:
            N.check(L.sar_runtime_image_download(f.rt, 0, 0, C.c_void_p(self.h_img.data_ptr()), sp))
        f.release_image(sp)
        if f.rank == 0:
            f.check_sync()      # the image was just consumed: a timed-out wait must not pass silently
