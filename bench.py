#!/usr/bin/env python
"""bench.py -- headline benchmark: FLAC encode MSamples/s at compression_level=5 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], per GPU): 256 independent 48 kHz stereo int16 streams of 10 s
(480 000 samples), blocksize 4096, level 5 -- 30 208 frames, 245.76 M channel-samples, 491.5 MB of PCM
(larger than the 126 MB L2, so consecutive timed steps cannot be served from cache).
One *step* = one pass of the whole encode path over the batch.  1 sample = 1 channel-sample.

  value      device-resident throughput: PCM already in HBM -> packed .flac images in HBM, CUDA events on
             the launching stream, max over ranks.
  e2e        same metric through the C-ABI call with HOST (pinned) buffers: H2D of the PCM and D2H of the
             packed bytes + index inside the timed region.
  roofline   dominant kernel's algorithmic bytes (PCM read once + FLAC bytes written once) / its CUDA-event time
             vs the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the reference's own libFLAC 1.4.3 (oracle/_ref) on the host cores, same PCM, bytes compared.

Multi-GPU: streams are independent, so ranks shard them with no data-path collective (weak scaling: every
rank encodes its own 256 streams); torch.distributed is used for the barrier and the max-over-ranks time.

Extra legs in the same JSON line (--quick skips them): `config_c5_shape` = BASELINE configs[4]'s per-GPU shard (4096 x 131072
stereo int16, level 5) on every rank (at --gpus 8 that is the 32768 streams of configs[4]); `config_c2` = BASELINE configs[2]
(4096 x 262144 mono 24-bit, 192 kHz, level 8) at --gpus 1; `decode_*` top-level keys = BASELINE configs[3].
Every rank byte-compares a sample of ITS OWN output with libFLAC in the same run (`cpu_baseline.ranks_checked`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_STREAMS = 256
N_SAMPLES = 480000
CHANNELS = 2
SAMPLE_RATE = 48000
BPS = 16
LEVEL = 5
BLOCKSIZE = 4096
METRIC = "encode_msamples_per_s_level5"
UNIT = "MSamples/s"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` on the bench batch, from the ncu --set full
    capture of THIS round's code (profiles/r02_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); None when
    the file has no entry for the kernel (a stale constant would be worse than no number)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f)
        e = t["kernels"].get(kernel)
        return (int(e["dram_read_bytes"]) + int(e["dram_write_bytes"]), t.get("source", "profiles/r02_traffic.json")) if e else (None, None)
    except Exception:
        return None, None


def workload_config():
    return {"workload": "StreamEncoder: 256 parallel 48kHz stereo int16 streams x 10 s, blocksize=4096, level=5 (BASELINE configs[1]) per GPU",
            "n_streams_per_gpu": N_STREAMS, "samples_per_stream": N_SAMPLES, "channels": CHANNELS,
            "bits_per_sample": BPS, "compression_level": LEVEL, "blocksize": BLOCKSIZE,
            "l2_policy": "inputs (491.5 MB/step) larger than L2 (126 MB); no explicit flush",
            "parallelism": "streams sharded across ranks, no data-path collective"}


def make_pcm(rank, n_streams=N_STREAMS):
    """Deterministic synthetic music-like PCM; 32 distinct seeds per rank, tiled with a
    per-stream circular shift + gain so every stream is different but generation stays fast."""
    from pyflac_b200.synth import music_like
    base = [music_like(N_SAMPLES, CHANNELS, SAMPLE_RATE, BPS, seed=1000 * rank + s) for s in range(32)]
    out = np.empty((n_streams, N_SAMPLES, CHANNELS), np.int16)
    for s in range(n_streams):
        b = base[s % 32]
        k = s // 32
        if k == 0:
            out[s] = b
        else:
            out[s] = np.roll(b, 7919 * k, axis=0)
            out[s] = (out[s].astype(np.int32) * (16 - k) // 16).astype(np.int16)
    return out


def make_pcm_short(rank, n_streams, n_samples):
    """Synthetic PCM for the decode leg (BASELINE configs[3] shape): 32 distinct music-like streams per rank, tiled with a
    circular shift + gain like make_pcm."""
    from pyflac_b200.synth import music_like
    base = [music_like(n_samples, CHANNELS, SAMPLE_RATE, BPS, seed=5000 + 1000 * rank + s) for s in range(32)]
    out = np.empty((n_streams, n_samples, CHANNELS), np.int16)
    for s in range(n_streams):
        b, k = base[s % 32], s // 32
        out[s] = b if k == 0 else (np.roll(b, 104729 * k % n_samples, axis=0).astype(np.int32) * (256 - k) // 256).astype(np.int16)
    return out


def make_pcm_c2(rank, n_streams, n_samples):
    """BASELINE configs[2]: mono 24-bit (int32 container) 192 kHz; 32 distinct music-like bases per rank, every stream a
    different circular shift + gain of one of them."""
    from pyflac_b200.synth import music_like
    base = [music_like(n_samples, 1, 192000, 24, seed=9000 + 1000 * rank + s)[:, 0] for s in range(32)]
    out = np.empty((n_streams, n_samples), np.int32)
    for s in range(n_streams):
        b, k = base[s % 32], s // 32
        out[s] = b if k == 0 else (np.roll(b, (15485863 * k) % n_samples).astype(np.int64) * (1024 - k) // 1024).astype(np.int32)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                if t0 <= t <= t1 + 0.2:
                    sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            if t0 <= t <= t1 + 0.2:
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_run(pcm, n_threads, keep_bytes=False, sample_rate=SAMPLE_RATE, bps=BPS, level=LEVEL):
    """libFLAC 1.4.3 (the binary pyFLAC bundles) over pthreads -- test/bench infrastructure from oracle/_ref."""
    if os.path.join(ROOT, "tests") not in sys.path:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _checkers as ck
    if not ck.ref_available():
        ck.build_checkers()
    return ck.ref_encode_mt(pcm, sample_rate, bps, level, BLOCKSIZE, n_threads, keep_bytes=keep_bytes)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation on the host cores, same config/metric."""
    if rank != 0:
        return
    hw = host_threads()
    n_sample = N_STREAMS
    pcm = make_pcm(0, n_sample)
    # pick the thread count libFLAC scales best with on this host (all hardware threads is not always it)
    cand = sorted({max(1, hw // 4), max(1, hw // 2), hw})
    probe = {t: cpu_reference_run(pcm, t)[0] for t in cand}
    threads = min(probe, key=probe.get)
    times = []
    for i in range(args.warmup + args.steps):
        dt, _, _ = cpu_reference_run(pcm, threads)
        if i >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    val = n_sample * N_SAMPLES * CHANNELS / t / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32/f64 (libFLAC)", "data": "synthetic", "config": workload_config(),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": f"{n_sample} of the {N_STREAMS} streams per step, {threads} pthreads, libFLAC 1.4.3 from oracle/_ref"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline encode only (profiling runs): no decode / config legs")
    ap.add_argument("--no-pipelined", action="store_true", help="skip the submit/collect leg (runs under ncu: a profiler that serialises kernels "
                    "starves md5_kernel of the arrival flags it waits for); e2e then reports the synchronous call")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from pyflac_b200 import _native as nat

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    pcm = make_pcm(rank)
    total_samples = pcm.size                       # channel-samples per rank per step
    pcm_bytes = pcm.nbytes
    h_pcm = torch.from_numpy(pcm.reshape(-1)).pin_memory()
    d_pcm = h_pcm.to(dev, non_blocking=False)
    stream_off = (np.arange(N_STREAMS, dtype=np.uint64) * np.uint64(N_SAMPLES * CHANNELS))
    stream_samples = np.full(N_STREAMS, N_SAMPLES, np.uint64)

    eng = nat.Engine(local_rank)
    # a real (non-legacy) torch stream: the engine launches on it and torch.cuda.Event times that same stream
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    eng.set_stream(work_stream.cuda_stream)
    eng.set_profiling(True)
    cfg = nat.Engine.make_config(SAMPLE_RATE, CHANNELS, BPS, LEVEL, BLOCKSIZE, container_bytes=2)

    def step_device():
        eng.encode_device(cfg, d_pcm.data_ptr(), d_pcm.numel(), stream_off, stream_samples)

    for _ in range(max(1, args.warmup)):           # (at least one: the first batch also sizes the result buffers below)
        step_device()
    torch.cuda.synchronize()
    res = eng.result()
    out_bytes = int(res.total_bytes)
    guard_hits = int(res.log_guard_hits)

    # ---------------- device-resident timing ----------------
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kt_acc = {}
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    eng.join()                                      # MD5 / finalize side streams of the in-flight steps end inside the timed region
    ev1.record()
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    if world > 1:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    kt_acc = eng.kernel_times()                    # per-kernel CUDA-event times of the last timed step
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    # ---------------- end to end (host buffers through the C ABI) ----------------
    arena_cap = pcm_bytes + (64 << 20)
    h_arena = torch.empty(arena_cap, dtype=torch.uint8).pin_memory()
    n_frames = int(res.n_frames)
    h_foff = np.zeros(n_frames, np.uint64)
    h_flen = np.zeros(n_frames, np.uint32)
    infos = (nat.StreamInfo * N_STREAMS)()
    import ctypes as C
    tot = C.c_uint64(0)
    L = nat.lib()

    def step_e2e():
        rc = L.flacb200_encode_batch_host(eng._h, C.byref(cfg), h_pcm.data_ptr(), h_pcm.numel(), N_STREAMS,
                                          stream_off.ctypes.data, stream_samples.ctypes.data, h_arena.data_ptr(), arena_cap,
                                          C.byref(tot), h_foff.ctypes.data, h_flen.ctypes.data, C.cast(infos, C.c_void_p))
        if rc != 0:
            raise RuntimeError(L.flacb200_last_error(eng._h).decode())

    for _ in range(max(1, args.warmup - 1)):
        step_e2e()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - e0
    hp = (C.c_double * 10)()
    L.flacb200_host_path_info(eng._h, hp, 10)
    host_path_ms = {"md5_workers_done": hp[0], "enqueued": hp[1], "kernels_done": hp[2], "d2h_done": hp[3], "md5_joined": hp[4], "total": hp[5],
                    "gpu_md5_done": hp[6], "streams_hashed_on_gpu": int(hp[7]), "host_md5_threads": int(hp[8]), "chunks": int(hp[9])}
    if world > 1:
        dist.barrier()

    # ---------------- end to end, batches in flight (flacb200_encode_host_submit / _collect) ----------------
    # The throughput form of the same public API: step i submits its batch (H2D from the pinned input inside the step) and
    # collects batch i-2 (its images, index and STREAMINFO digests read back to pinned host memory); every one of the K
    # batches is collected inside the timed region.  No call waits for a serial MD5 chain (md5_kernel, side stream), and the
    # host reads the PCM once (the DMA), not twice (DMA + host MD5) -- what limits many ranks sharing one host.
    DEPTH = 3
    p_arena = [torch.empty(arena_cap, dtype=torch.uint8).pin_memory() for _ in range(DEPTH)]
    p_foff = [np.zeros(n_frames, np.uint64) for _ in range(DEPTH)]
    p_flen = [np.zeros(n_frames, np.uint32) for _ in range(DEPTH)]
    p_info = [(nat.StreamInfo * N_STREAMS)() for _ in range(DEPTH)]
    p_tot = C.c_uint64(0)

    def submit(i):
        k = i % DEPTH
        tk = C.c_int(-1)
        rc = L.flacb200_encode_host_submit(eng._h, C.byref(cfg), h_pcm.data_ptr(), h_pcm.numel(), N_STREAMS, stream_off.ctypes.data,
                                           stream_samples.ctypes.data, p_arena[k].data_ptr(), arena_cap, p_foff[k].ctypes.data,
                                           p_flen[k].ctypes.data, C.cast(p_info[k], C.c_void_p), C.byref(tk))
        if rc != 0:
            raise RuntimeError(L.flacb200_last_error(eng._h).decode())
        return tk.value

    def collect(tk):
        if L.flacb200_encode_host_collect(eng._h, tk, C.byref(p_tot)) != 0:
            raise RuntimeError(L.flacb200_last_error(eng._h).decode())

    def run_pipelined(n):
        tickets = []
        for i in range(n):
            if len(tickets) == DEPTH:
                collect(tickets.pop(0))
            tickets.append(submit(i))
        while tickets:
            collect(tickets.pop(0))

    if args.no_pipelined:
        pipe_s, pipe_equal = e2e_s, True
    else:
        run_pipelined(max(DEPTH, args.warmup))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        p0 = time.perf_counter()
        run_pipelined(args.steps)
        torch.cuda.synchronize()
        pipe_s = time.perf_counter() - p0
        last = (args.steps - 1) % DEPTH
        pipe_equal = bool(p_tot.value == tot.value and np.array_equal(p_arena[last].numpy()[:tot.value], h_arena.numpy()[:tot.value])
                          and np.array_equal(p_foff[last], h_foff))
    del p_arena
    if world > 1:
        dist.barrier()

    # ---------------- every rank checks ITS OWN bytes against libFLAC, same run ----------------
    # (rank 0 additionally compares all of its 256 streams further down; here: a sample of >= 8 streams per rank, so that at
    #  N GPUs the whole job's output is covered, not just rank 0's)
    ranks_checked, ranks_equal = 0, True
    if not args.no_cpu_baseline:
        sel = sorted(set(int(v) for v in np.linspace(0, N_STREAMS - 1, 8)) | {(37 * rank + 5) % N_STREAMS})
        arena_chk = h_arena.numpy()
        _, _, ref_blobs = cpu_reference_run(np.ascontiguousarray(pcm[sel]), min(len(sel), max(1, host_threads() // max(world, 1))), keep_bytes=True)
        mine_ok = all(bytes(arena_chk[int(infos[s].byte_off): int(infos[s].byte_off + infos[s].byte_len)]) == ref_blobs[k].tobytes()
                      for k, s in enumerate(sel))
        del ref_blobs
        chk = torch.tensor([1.0 if mine_ok else 0.0, 1.0], dtype=torch.float64, device=dev)
        if world > 1:
            mn = chk[:1].clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            sm = chk[1:].clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ranks_equal, ranks_checked = bool(mn[0] > 0.5), int(round(float(sm[0])))
        else:
            ranks_equal, ranks_checked = bool(mine_ok), 1

    # ---------------- scatter / gather over NCCL (PCM born on rank 0's GPU) ----------------
    sg = None
    if world > 1 and not args.quick:
        try:
            from pyflac_b200.dist import scatter_streams, gather_packed

            class _DevMem:      # zero-copy torch view of the engine's device arena (CUDA array interface)
                def __init__(self, ptr, nbytes):
                    self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
            elems = N_SAMPLES * CHANNELS
            # setup (untimed): the root collects every rank's streams so that all PCM starts on one GPU
            mine = d_pcm.view(N_STREAMS, elems).view(torch.uint8)           # NCCL moves bytes, not int16
            parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
            dist.gather(mine, gather_list=parts, dst=0)
            full = torch.cat(parts, 0).view(torch.int16) if rank == 0 else None
            del parts

            def step_sg():
                blk, (lo, hi) = scatter_streams(full, world * N_STREAMS, elems, torch.int16, dev)
                k = hi - lo
                off = (np.arange(k, dtype=np.uint64) * np.uint64(elems))
                eng.encode_device(cfg, blk.data_ptr(), blk.numel(), off, np.full(k, N_SAMPLES, np.uint64))
                # frames never depend on the MD5 (a serial chain, ~20 ms for 10 s streams): gather the packed bytes as soon as
                # the frames are final and send the 16-byte digests after them
                r = eng.result(wait_md5=False)
                arena = torch.as_tensor(_DevMem(r.d_arena, int(r.total_bytes)), device=dev)
                bufs, sizes = gather_packed(arena, int(r.total_bytes), dev)
                if first is False:
                    send_digests(eng.fetch_md5_back(1))                            # the round before: its chain has finished by now
                return sizes

            def send_digests(d):
                dig = torch.from_numpy(d).to(dev)
                alld = [torch.empty_like(dig) for _ in range(world)] if rank == 0 else None
                dist.gather(dig, gather_list=alld, dst=0)
            first = True
            for _ in range(2):
                sizes = step_sg()
                first = False
            send_digests(eng.fetch_md5())
            torch.cuda.synchronize(); dist.barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nsg = max(3, args.steps // 4)
            first = True
            g0.record()
            for _ in range(nsg):
                sizes = step_sg()
                first = False
            send_digests(eng.fetch_md5())                                          # the last round's digests: the one wait of the run
            g1.record()
            torch.cuda.synchronize(); dist.barrier()
            tsg = torch.tensor([g0.elapsed_time(g1) / nsg], dtype=torch.float64, device=dev)
            dist.all_reduce(tsg, op=dist.ReduceOp.MAX)
            sg = {"value": world * total_samples / (float(tsg[0]) * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": float(tsg[0]),
                  "scatter_bytes_per_step": (world - 1) * pcm_bytes, "gather_bytes_per_step": int(sum(sizes[1:])),
                  "what": "all PCM resident on rank 0's GPU -> NCCL scatter of stream blocks -> encode on every rank -> NCCL gather of the packed "
                          "bytes to rank 0; the MD5 digests (16 B per stream) of round i travel during round i+1, the last round's inside the timed region"}
            del full
        except Exception as ex:                                   # never lose the main line to the optional mode
            sg = {"error": repr(ex)[:300]}
        torch.cuda.set_stream(work_stream)
        eng.set_stream(work_stream.cuda_stream)

    # ---------------- decode (second half of the metric), device-resident and host->host ----------------
    def decode_leg(blob_np, s_off, s_len, n_elems, expect, steps):
        """blob_np: pinned uint8 array (+16 bytes of padding); returns (device ms/step, e2e ms/step, kernel ms, ok)."""
        nbytes = int(blob_np.size - 16)
        d_flac = torch.from_numpy(blob_np).to(dev)
        for _ in range(2):
            eng.decode_device(d_flac.data_ptr(), nbytes, s_off, s_len, 2)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dv0, dv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dv0.record()
        for _ in range(steps):
            eng.decode_device(d_flac.data_ptr(), nbytes, s_off, s_len, 2)
        dv1.record()
        torch.cuda.synchronize()
        dev_ms = dv0.elapsed_time(dv1) / steps
        kt = eng.decode_kernel_times()
        ok = int(eng.decode_result().total_elems) == n_elems
        del d_flac
        h_out = torch.empty(n_elems, dtype=torch.int16).pin_memory()
        infos_d = [None]

        def step_e2e():
            infos_d[0] = eng.decode_host_pipelined(blob_np, s_off, s_len, h_out.numpy(), 2)[1]
        step_e2e()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        d0 = time.perf_counter()
        for _ in range(steps):
            step_e2e()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - d0) * 1e3 / steps
        ok = ok and bool(np.array_equal(h_out.numpy(), expect.reshape(-1))) and all(infos_d[0][s].status == 0 for s in range(len(s_off)))
        del h_out
        return dev_ms, e2e_ms, kt, ok

    def encode_leg(cfg_x, pcm_np, n_streams, n_samples, channels, steps, with_e2e):
        """One extra encode workload on this rank: device-resident ms/step (CUDA events) and, optionally, host->host ms/step."""
        flat = pcm_np.reshape(-1)
        h_x = torch.from_numpy(flat).pin_memory() if with_e2e else None
        d_x = (h_x if with_e2e else torch.from_numpy(flat)).to(dev)
        off_x = np.arange(n_streams, dtype=np.uint64) * np.uint64(n_samples * channels)
        smp_x = np.full(n_streams, n_samples, np.uint64)
        for _ in range(2):
            eng.encode_device(cfg_x, d_x.data_ptr(), d_x.numel(), off_x, smp_x)
        torch.cuda.synchronize()
        r = eng.result()
        if world > 1:
            dist.barrier()
        l0 = eng.launch_count
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(steps):
            eng.encode_device(cfg_x, d_x.data_ptr(), d_x.numel(), off_x, smp_x)
        eng.join()
        a1.record()
        torch.cuda.synchronize()
        out = {"ms_dev": a0.elapsed_time(a1) / steps, "kernel_ms": eng.kernel_times(), "out_bytes": int(r.total_bytes), "n_frames": int(r.n_frames),
               "launches_per_step": (eng.launch_count - l0) // steps, "guard_hits": int(r.log_guard_hits), "d_pcm": d_x, "off": off_x, "smp": smp_x, "ms_e2e": None}
        if with_e2e:
            cap = flat.nbytes + (64 << 20)
            h_ar = torch.empty(cap, dtype=torch.uint8).pin_memory()
            foff = np.zeros(out["n_frames"], np.uint64); flen = np.zeros(out["n_frames"], np.uint32)
            inf = (nat.StreamInfo * n_streams)()
            totx = C.c_uint64(0)

            def st():
                rc = L.flacb200_encode_batch_host(eng._h, C.byref(cfg_x), h_x.data_ptr(), h_x.numel(), n_streams, off_x.ctypes.data, smp_x.ctypes.data,
                                                  h_ar.data_ptr(), cap, C.byref(totx), foff.ctypes.data, flen.ctypes.data, C.cast(inf, C.c_void_p))
                if rc != 0:
                    raise RuntimeError(L.flacb200_last_error(eng._h).decode())
            st()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                st()
            torch.cuda.synchronize()
            out["ms_e2e"] = (time.perf_counter() - t0) * 1e3 / steps
            out["e2e_d2h_bytes"] = int(totx.value) + out["n_frames"] * 12 + n_streams * 56
            del h_ar
        return out

    extra_t = [0.0] * 10          # times to max-reduce: dec dev, dec e2e, rt dev, rt e2e, c5 dev, c5 e2e, c2 dev
    dec = rt = c5 = c2 = None
    if not args.quick:
        # (a) round trip: the 256 streams the encode step just produced
        arena_np = h_arena.numpy()
        total_flac = int(tot.value)
        s_off = np.array([infos[s].byte_off for s in range(N_STREAMS)], np.uint64)
        s_len = np.array([infos[s].byte_len for s in range(N_STREAMS)], np.uint64)
        rt_dev_ms, rt_e2e_ms, rt_kt, rt_ok = decode_leg(arena_np[:total_flac + 16], s_off, s_len, total_samples, pcm, args.steps)
        rt = {"kt": rt_kt, "ok": rt_ok}
        extra_t[2], extra_t[3] = rt_dev_ms, rt_e2e_ms

        # (b) BASELINE configs[4]'s per-GPU shard: 4096 streams x 131072 stereo int16, level 5 (at --gpus 8: 32768 streams)
        D_STREAMS, D_SAMPLES = 4096, 131072
        pcm4 = make_pcm_short(rank, D_STREAMS, D_SAMPLES)
        xsteps = max(3, args.steps // 4)
        c5 = encode_leg(cfg, pcm4, D_STREAMS, D_SAMPLES, CHANNELS, xsteps, with_e2e=True)
        extra_t[4], extra_t[5] = c5["ms_dev"], c5["ms_e2e"]

        # (c) BASELINE configs[3]: those 4096 .flac streams -> int16 PCM
        eng.encode_device(cfg, c5["d_pcm"].data_ptr(), c5["d_pcm"].numel(), c5["off"], c5["smp"])
        enc4 = eng.fetch()
        c5.pop("d_pcm")
        total4 = int(enc4["total_bytes"])
        h_blob4 = torch.empty(total4 + 16, dtype=torch.uint8).pin_memory()
        h_blob4.numpy()[:total4] = enc4["arena"][:total4]
        h_blob4.numpy()[total4:] = 0
        s_off4 = np.array([si.byte_off for si in enc4["streams"]], np.uint64)
        s_len4 = np.array([si.byte_len for si in enc4["streams"]], np.uint64)
        del enc4
        dsteps = max(3, args.steps // 4)
        dec_ms_step, dec_e2e_ms_step, dec_kt, dec_ok = decode_leg(h_blob4.numpy(), s_off4, s_len4, pcm4.size, pcm4, dsteps)
        dec = {"kt": dec_kt, "ok": dec_ok, "samples": pcm4.size, "bytes": total4 + pcm4.nbytes, "flac_bytes": total4, "pcm_bytes": int(pcm4.nbytes), "steps": dsteps}
        extra_t[0], extra_t[1] = dec_ms_step, dec_e2e_ms_step
        c5_cpu = None
        if rank == 0 and not args.no_cpu_baseline:
            import _checkers as ck
            thr = host_threads()
            ddt = min(ck.ref_decode_mt(h_blob4.numpy(), s_off4, s_len4, thr)[0] for _ in range(2))
            dec["cpu_value"], dec["cpu_cores"] = pcm4.size / ddt / 1e6, thr
            # libFLAC on a bounded sample of the c5 shape (512 of the 4096 streams) + byte equality of those streams
            nsm = 512
            dtc, _, blobs = cpu_reference_run(pcm4[:nsm], thr, keep_bytes=True)
            eng.encode_host(cfg, pcm4[:nsm].reshape(-1), np.arange(nsm, dtype=np.uint64) * np.uint64(D_SAMPLES * CHANNELS), np.full(nsm, D_SAMPLES, np.uint64))
            o5 = eng.fetch()
            eq5 = all(o5["arena"][int(si.byte_off): int(si.byte_off + si.byte_len)].tobytes() == blobs[k].tobytes() for k, si in enumerate(o5["streams"]))
            c5_cpu = {"value": nsm * D_SAMPLES * CHANNELS / dtc / 1e6, "unit": UNIT, "cores": thr, "kind": "reference",
                      "sample": f"{nsm} of the 4096 streams, one FLAC__StreamEncoder per pthread", "bytes_identical_to_gpu": bool(eq5)}
            del blobs, o5
        del h_blob4, pcm4

        # (d) BASELINE configs[2]: 4096 streams x 262144 mono 24-bit (int32 container), 192 kHz, level 8 -- 1 GPU only
        if world == 1:
            C2_STREAMS, C2_SAMPLES = 4096, 262144
            pcm2 = make_pcm_c2(rank, C2_STREAMS, C2_SAMPLES)
            cfg2 = nat.Engine.make_config(192000, 1, 24, 8, BLOCKSIZE, container_bytes=4)
            c2 = encode_leg(cfg2, pcm2, C2_STREAMS, C2_SAMPLES, 1, 3, with_e2e=False)
            c2.pop("d_pcm")
            c2["pcm_bytes"], c2["samples"] = int(pcm2.nbytes), int(pcm2.size)
            extra_t[6] = c2["ms_dev"]
            if not args.no_cpu_baseline:
                thr = host_threads()
                nsm = 256
                dtc, _, blobs = cpu_reference_run(pcm2[:nsm, :, None], thr, keep_bytes=True, sample_rate=192000, bps=24, level=8)
                eng.encode_host(cfg2, pcm2[:nsm].reshape(-1), np.arange(nsm, dtype=np.uint64) * np.uint64(C2_SAMPLES), np.full(nsm, C2_SAMPLES, np.uint64))
                o2 = eng.fetch()
                eq2 = all(o2["arena"][int(si.byte_off): int(si.byte_off + si.byte_len)].tobytes() == blobs[k].tobytes() for k, si in enumerate(o2["streams"]))
                c2["cpu"] = {"value": nsm * C2_SAMPLES / dtc / 1e6, "unit": UNIT, "cores": thr, "kind": "reference",
                             "sample": f"{nsm} of the 4096 streams, libFLAC 1.4.3 with set_bits_per_sample(24), level 8", "bytes_identical_to_gpu": bool(eq2)}
                del blobs, o2
            del pcm2

    # ---------------- reduce over ranks ----------------
    t = torch.tensor([ms_total, e2e_s * 1e3, pipe_s * 1e3, 0.0 if pipe_equal else 1.0] + extra_t, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max, e2e_ms_max, pipe_ms_max, pipe_bad = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    xt = [float(v) for v in t[4:]]

    if rank == 0:
        ms_per_step = ms_total_max / args.steps
        value = world * total_samples / (ms_per_step * 1e-3) / 1e6
        sync_val = world * total_samples / (e2e_ms_max / args.steps * 1e-3) / 1e6
        e2e_val = world * total_samples / (pipe_ms_max / args.steps * 1e-3) / 1e6
        peak, peak_src = hbm_peak()
        # dominant kernel = the longest one on the encode stream (the step's critical path).  md5_kernel runs on a side
        # stream under the next batches (a 256-thread serial chain, pure latency) and is listed in kernel_ms / roofline_md5.
        dom = max([k for k in ("fused", "autoc", "analyze", "pack") if k in kt_acc], key=lambda k: kt_acc.get(k, 0.0))
        alg_bytes = pcm_bytes + out_bytes
        achieved = alg_bytes / (kt_acc[dom] * 1e-3) / 1e9
        dom_kernel = ("fused_" if kt_acc.get("path") == "tma" and dom != "autoc" else "") + dom + "_kernel"
        traffic, traffic_src = ncu_traffic(dom_kernel)
        dec_traffic, dec_traffic_src = ncu_traffic("dec_frame_kernel")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32 (+f32 window, f64 autocorrelation/Levinson, as libFLAC)", "data": "synthetic",
            "config": workload_config(),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": pcm_bytes,
                    "d2h_bytes_per_step": out_bytes + n_frames * 12 + N_STREAMS * 56, "ms_per_step": pipe_ms_max / args.steps,
                    "mode": ("flacb200_encode_batch_host, one synchronous call per step (--no-pipelined)" if args.no_pipelined else
                             "flacb200_encode_host_submit / _collect, 3 batches in flight: step i copies its PCM in from pinned host memory and "
                             "reads batch i-2's images + index + STREAMINFO digests back; all K batches collected inside the timed region"),
                    "bytes_identical_to_sync_call": pipe_bad == 0.0,
                    # the step is the H2D copy: what the host's PCIe / memory fabric gives each GPU when all N ranks copy at once
                    "link_GBps": {"h2d_per_gpu": pcm_bytes / (pipe_ms_max / args.steps * 1e-3) / 1e9,
                                  "d2h_per_gpu": out_bytes / (pipe_ms_max / args.steps * 1e-3) / 1e9,
                                  "h2d_plus_d2h_all_gpus": world * (pcm_bytes + out_bytes) / (pipe_ms_max / args.steps * 1e-3) / 1e9},
                    "sync_call": {"value": sync_val, "unit": UNIT, "ms_per_step": e2e_ms_max / args.steps,
                                  "what": "flacb200_encode_batch_host: one synchronous call per step, complete results (incl. MD5) on return",
                                  "last_call_breakdown_ms": host_path_ms}},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "step_frac": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                         "kernel_ms": kt_acc,
                         "kernel_ms_note": "CUDA events on the launching streams around each kernel of the LAST timed step (steady state; the step "
                                           "time above is the average over all timed steps)"},
            "roofline_md5": {"bound": "hbm", "kernel": "md5_kernel (side stream, overlapped with the next batches)",
                             "achieved": pcm_bytes / (max(kt_acc.get("md5", 0.0), 1e-6) * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": pcm_bytes / (max(kt_acc.get("md5", 0.0), 1e-6) * 1e-3) / 1e9 / peak,
                             "algorithmic_bytes_per_launch": pcm_bytes},
            "scatter_gather": sg,
            "frames_per_step": n_frames * world, "frames_per_s": n_frames * world / (ms_per_step * 1e-3),
            "x_realtime": value * 1e6 / (SAMPLE_RATE * CHANNELS), "x_realtime_e2e": e2e_val * 1e6 / (SAMPLE_RATE * CHANNELS),
            "compressed_bytes_per_step": out_bytes, "ratio": out_bytes / pcm_bytes,
            "log_guard_hits": guard_hits,
        }
        if dec is not None:
            dec_ms_max, dec_e2e_ms_max, rt_dev_max, rt_e2e_max, c5_dev_max, c5_e2e_max, c2_dev_max = xt[0], xt[1], xt[2], xt[3], xt[4], xt[5], xt[6]
            dval = world * dec["samples"] / (dec_ms_max * 1e-3) / 1e6
            dval_e2e = world * dec["samples"] / (dec_e2e_ms_max * 1e-3) / 1e6
            # BASELINE configs[3] lifted to the top level (second half of the metric)
            line["decode_metric"] = "decode_msamples_per_s"
            line["decode_value"] = dval
            line["decode_e2e_value"] = dval_e2e
            line["decode_ms_per_step"] = dec_ms_max
            line["decode"] = {"metric": "decode_msamples_per_s", "unit": UNIT, "value": dval, "e2e_value": dval_e2e,
                              "ms_per_step": dec_ms_max, "e2e_ms_per_step": dec_e2e_ms_max, "steps": dec["steps"],
                              "workload": "StreamDecoder: 4096 parallel .flac streams (131072 stereo int16 samples each, level 5) -> int16 PCM (BASELINE configs[3]) per GPU",
                              "h2d_bytes_per_step": dec["flac_bytes"], "d2h_bytes_per_step": dec["pcm_bytes"],
                              "pcm_identical_to_input": bool(dec["ok"]), "kernel_ms": dec["kt"],
                              "cpu_baseline": ({"value": dec["cpu_value"], "unit": UNIT, "cores": dec["cpu_cores"], "kind": "reference",
                                                "sample": "all 4096 streams, one FLAC__StreamDecoder per pthread"} if "cpu_value" in dec else None),
                              "roofline": {"bound": "hbm", "kernel": "dec_frame_kernel",
                                           "achieved": dec["bytes"] / (max(dec["kt"].get("frame_decode", 0.0) - dec["kt"].get("crc16", 0.0), 1e-6) * 1e-3) / 1e9,
                                           "peak": peak, "unit": "GB/s",
                                           "frac": dec["bytes"] / (max(dec["kt"].get("frame_decode", 0.0) - dec["kt"].get("crc16", 0.0), 1e-6) * 1e-3) / 1e9 / peak,
                                           "kernel_ms_note": "frame_decode = dec_frame_kernel + dec_crc_kernel (crc16 is the latter's share); the roofline line is dec_frame_kernel alone",
                                           "step_frac": dec["bytes"] / (dec_ms_max * 1e-3) / 1e9 / peak,
                                           "traffic": dec_traffic, "traffic_source": dec_traffic_src,
                                           "algorithmic_bytes_per_launch": dec["bytes"]}}
            line["decode_roundtrip_256"] = {"value": world * total_samples / (rt_dev_max * 1e-3) / 1e6, "e2e_value": world * total_samples / (rt_e2e_max * 1e-3) / 1e6,
                                            "unit": UNIT, "ms_per_step": rt_dev_max, "e2e_ms_per_step": rt_e2e_max, "kernel_ms": rt["kt"],
                                            "workload": "the 256 streams produced by the encode step (libFLAC-identical bytes) -> int16 PCM; a launch cannot "
                                                        "finish faster than one frame's serial decode, so this small batch is latency bound",
                                            "pcm_identical_to_input": bool(rt["ok"])}
            c5_samples = 4096 * 131072 * CHANNELS
            c5_alg = c5_samples * 2 + c5["out_bytes"]
            c5k = c5["kernel_ms"]
            c5dom = max([k for k in ("fused", "autoc", "analyze", "pack") if k in c5k], key=lambda k: c5k.get(k, 0.0))
            line["config_c5_shape"] = {
                "workload": f"BASELINE configs[4]: {4096 * world} streams x 131072 stereo int16 samples, level 5, blocksize 4096 ({4096} per GPU, {world} GPU(s))",
                "metric": METRIC, "unit": UNIT, "value": world * c5_samples / (c5_dev_max * 1e-3) / 1e6, "ms_per_step": c5_dev_max,
                "e2e": {"value": world * c5_samples / (c5_e2e_max * 1e-3) / 1e6, "ms_per_step": c5_e2e_max, "h2d_bytes_per_step": c5_samples * 2,
                        "d2h_bytes_per_step": c5.get("e2e_d2h_bytes")},
                "n_streams_total": 4096 * world, "frames_per_step": c5["n_frames"] * world, "gpu_launches_per_step": c5["launches_per_step"],
                "roofline": {"bound": "hbm", "kernel": ("fused_" if c5k.get("path") == "tma" and c5dom != "autoc" else "") + c5dom + "_kernel",
                             "achieved": c5_alg / (c5k[c5dom] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": c5_alg / (c5k[c5dom] * 1e-3) / 1e9 / peak,
                             "algorithmic_bytes_per_launch": c5_alg, "kernel_ms": c5k, "traffic": None},
                "cpu_baseline": c5_cpu, "log_guard_hits": c5["guard_hits"]}
            if c2 is not None:
                c2_alg = c2["pcm_bytes"] + c2["out_bytes"]
                c2k = c2["kernel_ms"]
                c2dom = max([k for k in ("fused", "autoc", "analyze", "pack") if k in c2k], key=lambda k: c2k.get(k, 0.0))
                line["config_c2"] = {
                    "workload": "BASELINE configs[2]: StreamEncoder, 4096 streams x 262144 mono 24-bit samples (int32 container), 192 kHz, level 8, blocksize 4096, 1 GPU",
                    "metric": "encode_msamples_per_s_level8_24bit", "unit": UNIT, "value": c2["samples"] / (c2_dev_max * 1e-3) / 1e6, "ms_per_step": c2_dev_max,
                    "frames_per_step": c2["n_frames"], "gpu_launches_per_step": c2["launches_per_step"], "ratio": c2["out_bytes"] / c2["pcm_bytes"],
                    "roofline": {"bound": "hbm", "kernel": c2dom + "_kernel", "achieved": c2_alg / (c2k[c2dom] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": c2_alg / (c2k[c2dom] * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": c2_alg, "kernel_ms": c2k, "traffic": None},
                    "cpu_baseline": c2.get("cpu"), "log_guard_hits": c2["guard_hits"]}
        if not args.no_cpu_baseline:
            threads = host_threads()
            n_sample = N_STREAMS
            _, _, blobs = cpu_reference_run(pcm[:n_sample], threads, keep_bytes=True)
            # same-run byte equality of every stream of rank 0 against the GPU output of the last e2e step
            arena = h_arena.numpy()
            equal = all(bytes(arena[int(infos[s].byte_off): int(infos[s].byte_off + infos[s].byte_len)]) == blobs[s].tobytes()
                        for s in range(n_sample))
            del blobs
            # same procedure as --impl reference: pick the thread count libFLAC scales best with, then average timed runs
            cand = sorted({max(1, threads // 4), max(1, threads // 2), threads})
            probe = {tt: cpu_reference_run(pcm[:n_sample], tt)[0] for tt in cand}
            tbest = min(probe, key=probe.get)
            dts = [cpu_reference_run(pcm[:n_sample], tbest)[0] for _ in range(3)]
            one = min(cpu_reference_run(pcm[:8], 1)[0] for _ in range(2))
            line["cpu_baseline"] = {"value": n_sample * N_SAMPLES * CHANNELS / float(np.mean(dts)) / 1e6, "unit": UNIT, "cores": tbest, "kind": "reference",
                                    "sample": f"all {n_sample} streams of the step, one FLAC__StreamEncoder per pthread, "
                                              f"libFLAC 1.4.3 from oracle/_ref; thread count picked by a probe of {cand}, mean of 3 runs",
                                    "one_thread_value": 8 * N_SAMPLES * CHANNELS / one / 1e6,
                                    "host_threads": threads, "bytes_identical_to_gpu": bool(equal and ranks_equal),
                                    "ranks_checked": ranks_checked,
                                    "streams_checked": f"rank 0: all {n_sample}; every rank: 9 of its own {N_STREAMS}"}
            if dec is not None and "cpu_value" in dec:
                line["cpu_baseline"]["decode_value"], line["cpu_baseline"]["decode_cores"] = dec["cpu_value"], dec["cpu_cores"]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
