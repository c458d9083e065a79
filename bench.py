#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on B200: KITTI-shaped ORB extraction (+ stereo matching) in
frames/s at nFeatures = 2000, and brute-force Hamming matching in Gpairs/s.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch of synthetic KITTI-shaped stereo pairs
(BASELINE.json configs[1]: 1241x376, nFeatures 2000, extraction of both images +
Frame::ComputeStereoMatches).  `value` is whole-job frames/s with the batch resident in HBM;
`e2e` is the same work through the host-buffer C-ABI calls with the H2D / D2H copies inside the
timed region.  The reference arm (and the cpu_baseline leg of the own arm) times the reference's OWN ORBextractor.cc
— compiled from /root/reference against an OpenCV API stand-in into oracle/_ref, see oracle/ref_shim — plus the restated
Frame::ComputeStereoMatches (pinned bit-exact to the reference's Frame.cc) on all host cores, one frame pair per thread,
on a bounded sample per step; where oracle/_ref is absent it falls back to the CPU oracle port (oracle/oracle_cpu.cpp).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

W, H, NFEATURES = 1241, 376, 2000
PARAMS = (NFEATURES, 1.2, 8, 20, 7)          # Examples/Stereo/KITTI00-02.yaml:38-51
BF, FX = 386.1448, 718.856                   # Examples/Stereo/KITTI00-02.yaml:8,25
LEVEL_PIXELS = 1444097                       # sum of the 8 level sizes of a 1241x376 image
# bytes per image = every pyramid pixel read once + every derived-level pixel written once
# + 60 B per keypoint = 2.542 MB at ~2006 keypoints
BYTES_PER_IMAGE = LEVEL_PIXELS + (LEVEL_PIXELS - W * H) + 60 * 2006
HAMMING_NQ, HAMMING_BANK = 2000, 1_000_000   # BASELINE.json configs[3]


def bind_to_gpu_numa_node(gpu_index):
    """Pins this rank to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is allocated, so that
    first-touch places the staging buffers next to the GPU's PCIe root (the e2e leg moves ~190 MB per step per rank)."""
    try:
        bus = subprocess.check_output(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"], text=True).strip()
        bus = bus.lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def log(msg):
    if os.environ.get("BENCH_VERBOSE"):
        print("[bench rank %s] %s" % (os.environ.get("RANK", "0"), msg), file=sys.stderr, flush=True)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [l.strip().split(", ") for l in open(self.path) if l.strip()]
            os.unlink(self.path)
        except Exception:
            rows = []
        sm, reasons, mx = [], set(), None
        for r in rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_reference():
    """(stereo_batch function, kind, description) of the CPU arm: the reference's own ORBextractor.cc where the checker
    built from /root/reference is present (oracle/_ref, travels with the snapshot), else the oracle port."""
    oracle_dir = os.path.join(ROOT, "oracle")          # the ONLY place this file reaches into oracle/: the CPU arm
    if oracle_dir not in sys.path:
        sys.path.insert(0, oracle_dir)
    try:
        import ref
        if ref.available():
            ref.lib()
            return ref.stereo_batch, "reference", ("reference's own ORBextractor.cc (oracle/_ref, OpenCV primitives = cv2-pinned restatements) "
                                                   "+ restated Frame::ComputeStereoMatches, g++ -O3, no -march")
    except Exception:
        pass
    import oracle
    return oracle.stereo_batch, "port", "CPU oracle port of the reference path (oracle/oracle_cpu.cpp), g++ -O3, no -march"


def make_inputs(pairs, seed):
    from refactored_orb_slam2_b200 import synthetic          # input generator only; nothing under oracle/ is touched by the own arm
    left, right = synthetic.frame_batch(pairs, H, W, seed, stereo=True)
    return np.ascontiguousarray(np.concatenate([left, right]))        # [2*pairs, H, W]: left images then right images


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return 0
    stereo_batch, kind, host = cpu_reference()
    cores = os.cpu_count() or 1
    pairs = max(16, 12 * cores) if args.ref_pairs <= 0 else args.ref_pairs      # 12 pairs per thread: enough for the threads to balance (~0.8 s per step)
    imgs = make_inputs(pairs, 100)
    left, right = imgs[:pairs], imgs[pairs:]
    mb = np.float32(np.float32(BF) / np.float32(FX))
    for _ in range(args.warmup):
        stereo_batch(PARAMS, left[:cores], right[:cores], BF, mb, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        stereo_batch(PARAMS, left, right, BF, mb, cores)
    dt = time.perf_counter() - t0
    fps = 2 * pairs * args.steps / dt
    sample = "%d KITTI-shaped stereo pairs per step, extraction L+R + ComputeStereoMatches, one pair per thread" % pairs
    line = {
        "impl": "reference", "metric": "KITTI ORB extract frames/s @2000 kp", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "KITTI 1241x376 stereo pairs, nFeatures=2000, extract L+R + ComputeStereoMatches",
                   "pairs_per_step": pairs, "host": host},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------- own arm
def run_own(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import refactored_orb_slam2_b200 as pkg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    # stdout carries exactly ONE JSON line: anything libraries print there (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = pkg.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pairs = args.pairs
    nimg = 2 * pairs
    log("world %d local_rank %d: generating %d pairs" % (world, local_rank, pairs))
    host = torch.from_numpy(make_inputs(pairs, 100 + rank)).pin_memory()
    dev = host.cuda()
    mb = float(np.float32(np.float32(BF) / np.float32(FX)))

    ex = pkg.ORBextractor(*PARAMS, max_width=W, max_height=H, max_batch=nimg, device=local_rank)
    stream = torch.cuda.ExternalStream(ex.stream, device=torch.device("cuda", local_rank))

    def step_device():
        ex.extract_batch_device(dev.data_ptr(), nimg, W, H)
        pkg._lib.check(lib.orbgpu_stereo_match_device(ex.handle, ex.handle, 0, pairs, pairs, BF, mb))

    log("handle ready")
    # ---- device-resident throughput (`value`)
    visible = os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")
    smi_index = int(visible[local_rank]) if local_rank < len(visible) and visible[local_rank].strip().isdigit() else local_rank
    sampler = ClockSampler(smi_index)
    if rank == 0:
        sampler.start()                            # sampled from warm-up to the end of the e2e loop (the GPU is loaded throughout)
    t_warm = time.perf_counter()
    for _ in range(args.warmup):
        step_device()
    ex.synchronize()
    while time.perf_counter() - t_warm < 0.5:      # keep the clocks up until the sampler has ticked a few times
        step_device()
        ex.synchronize()
    lib.orbgpu_launch_count(1)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    launches = int(lib.orbgpu_launch_count(0))
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ex.synchronize()                               # raises if a device buffer overflowed
    # per-stage device times from a separate, untimed pass: the stage timers serialise the blur, which the timed loop runs on a
    # side stream next to FAST + octree (so the stages add up to slightly more than ms_per_step)
    ex.set_profiling(True)
    ex.stage_times()
    for _ in range(max(3, min(args.steps, 8))):
        step_device()
    stages = ex.stage_times()
    ex.set_profiling(False)
    ms_step = ms_total / args.steps
    fps = world * nimg * args.steps / (ms_total * 1e-3)

    log("device loop done: %.3f ms/step" % ms_step)
    kps, desc, counts = ex.download_results()
    mean_kp = float(counts.mean())

    # ---- end to end through the host-buffer API (`e2e`): pinned host images in, keypoints /
    # descriptors / mvuRight / mvDepth out, every step
    cap = ex.capacity
    out_k = torch.empty((nimg, cap, 28), dtype=torch.uint8).pin_memory()
    out_d = torch.empty((nimg, cap, 32), dtype=torch.uint8).pin_memory()
    out_c = torch.empty(nimg, dtype=torch.int32).pin_memory()
    out_u = torch.empty((pairs, cap), dtype=torch.float32).pin_memory()
    out_z = torch.empty((pairs, cap), dtype=torch.float32).pin_memory()
    vp = ctypes.c_void_p

    def step_e2e():
        pkg._lib.check(lib.orbgpu_extract_batch(ex.handle, vp(host.data_ptr()), nimg, W, H, W, W * H, vp(out_k.data_ptr()),
                                                vp(out_d.data_ptr()), cap, vp(out_c.data_ptr())))
        pkg._lib.check(lib.orbgpu_stereo_match(ex.handle, ex.handle, 0, pairs, pairs, BF, mb, vp(out_u.data_ptr()),
                                               vp(out_z.data_ptr()), cap))

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    dt_sync = max_over_ranks(time.perf_counter() - t0)
    # the same calls, two batches in flight (pkg.StereoBatchPipeline: two handles, split async / wait entry points, one host thread), so that one
    # batch uploads while the other computes; every batch is copied in from pinned host memory and its results copied out
    del out_k, out_d, out_c, out_u, out_z
    depth = max(2, env_int("BENCH_PIPELINE_DEPTH", 2))
    pipe = pkg.StereoBatchPipeline(PARAMS, W, H, pairs, BF, mb, depth=depth, device=local_rank)
    for _ in range(depth):
        pipe.submit(host)
    for _ in range(depth):
        pipe.result()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        if i >= depth:
            pipe.result()          # batch i - depth: frees its handle; the batches after it are still in flight while batch i is enqueued
        pipe.submit(host)
    for _ in range(min(depth, args.steps) - 1):
        pipe.result()
    last = pipe.result()
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0)
    kmax = int(last[2].max())
    pipe.close()
    # ---- the reference's own calling pattern: ONE stereo pair per call (configs[1] as Tracking::GrabImageStereo runs it);
    # informational, the small-batch path replays a CUDA graph of the 13-kernel extraction sequence
    latency_ms = None
    if rank == 0 and not args.no_latency:
        ex1 = pkg.ORBextractor(*PARAMS, max_width=W, max_height=H, max_batch=2, device=local_rank)
        cap1 = ex1.capacity
        pair = torch.stack([host[0], host[pairs]]).pin_memory()
        o = [torch.empty((2, cap1, 28), dtype=torch.uint8).pin_memory(), torch.empty((2, cap1, 32), dtype=torch.uint8).pin_memory(),
             torch.empty(2, dtype=torch.int32).pin_memory(), torch.empty((1, cap1), dtype=torch.float32).pin_memory(),
             torch.empty((1, cap1), dtype=torch.float32).pin_memory()]
        ts = []
        for i in range(60):
            t1 = time.perf_counter()
            pkg._lib.check(lib.orbgpu_extract_batch(ex1.handle, vp(pair.data_ptr()), 2, W, H, W, W * H, vp(o[0].data_ptr()), vp(o[1].data_ptr()), cap1,
                                                    vp(o[2].data_ptr())))
            pkg._lib.check(lib.orbgpu_stereo_match(ex1.handle, ex1.handle, 0, 1, 1, BF, mb, vp(o[3].data_ptr()), vp(o[4].data_ptr()), cap1))
            if i >= 10:
                ts.append(time.perf_counter() - t1)
        latency_ms = float(np.median(ts) * 1e3)
        del ex1
    e2e_fps = world * nimg * args.steps / dt
    clocks = sampler.stop() if rank == 0 else None
    h2d = nimg * W * H
    d2h = nimg * 4 + nimg * cap * (28 + 32) + 2 * pairs * cap * 4      # whole rows are copied (counts live on the device); max count per image is reported in config

    log("e2e loop done")
    # ---- Hamming matching, BASELINE.json configs[3]: 2000 queries x 1M-row bank, row-sharded over the ranks,
    # all-gather of the per-shard {d1, d2, idx} records, merge + ratio test on every rank
    hamming = None
    if not args.no_hamming:
        lo, hi = rank * HAMMING_BANK // world, (rank + 1) * HAMMING_BANK // world
        shard = np.random.default_rng(4 + 1000 * rank).integers(0, 256, size=(hi - lo, 32), dtype=np.uint8)
        q = np.random.default_rng(5).integers(0, 256, size=(HAMMING_NQ, 32), dtype=np.uint8)      # same queries on every rank
        flips = np.random.default_rng(6).integers(0, 2, size=(HAMMING_NQ, 32), dtype=np.uint8)   # <= 32 flipped bits
        mine = np.arange(rank, HAMMING_NQ, 2 * world)                                             # half the queries have a near match
        shard[(mine * 7919) % (hi - lo)] = q[mine] ^ flips[mine]
        d_shard = torch.from_numpy(shard).cuda()
        d_q = torch.from_numpy(q).cuda()
        bank = pkg.DescriptorBank(device_ptr=d_shard.data_ptr(), rows=hi - lo, index_base=lo, device=local_rank, keepalive=d_shard)
        local = torch.empty((HAMMING_NQ, 16), dtype=torch.uint8, device="cuda")
        gathered = torch.empty((world * HAMMING_NQ, 16), dtype=torch.uint8, device="cuda")   # rank-major
        merged = torch.empty((HAMMING_NQ, 16), dtype=torch.uint8, device="cuda")
        match = torch.empty(HAMMING_NQ, dtype=torch.int64, device="cuda")
        cur = torch.cuda.Stream()          # an explicit side stream: a NULL stream argument means "the bank's own stream"
        torch.cuda.set_stream(cur)

        use_peer = world > 1 and args.hamming_exchange == "peer"
        if use_peer:
            from refactored_orb_slam2_b200.distributed import PeerExchange
            xchg = PeerExchange(rank, world, HAMMING_NQ, local_rank)      # IPC handles all-gathered once; no collective afterwards

        def step_hamming():
            if use_peer:      # scan -> records stored into every peer's window over NVLink -> wait on arrival flags + merge
                pkg._lib.check(lib.orbgpu_hamming_top2_sharded_device(bank._h, xchg._h, vp(d_q.data_ptr()), HAMMING_NQ, vp(merged.data_ptr()), 0.7, 50,
                                                                      vp(match.data_ptr()), vp(cur.cuda_stream)))
                return
            bank.top2_device(d_q.data_ptr(), HAMMING_NQ, local.data_ptr(), stream=cur.cuda_stream)
            if world > 1:
                dist.all_gather_into_tensor(gathered, local)
                src = gathered
            else:
                src = local
            pkg.merge_top2_device(src.data_ptr(), world, HAMMING_NQ, merged.data_ptr(), 0.7, 50, match.data_ptr(),
                                  device=local_rank, stream=cur.cuda_stream)

        hsteps = max(args.steps, 10)
        for _ in range(max(3, args.warmup)):
            step_hamming()
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record(cur)
        for _ in range(hsteps):
            step_hamming()
        h1.record(cur)
        barrier()
        hms = max_over_ranks(h0.elapsed_time(h1)) / hsteps
        if use_peer:
            xchg.check()                   # a peer whose records never arrived would have been recorded on the device
        gpairs = HAMMING_NQ * HAMMING_BANK / (hms * 1e-3) / 1e9
        popc, lop3 = ctypes.c_double(), ctypes.c_double()
        pkg._lib.check(lib.orbgpu_measure_popc_peak(local_rank, ctypes.byref(popc), ctypes.byref(lop3)))
        # integer-pipe roofline: algorithmic work = 8 POPC per descriptor pair
        peak_gpairs = world * popc.value / 8 / 1e9
        hamming = {"metric": "Hamming match Gpairs/s (2000 queries x 1M-row bank, top-2 + ratio 0.7 / TH_LOW 50)",
                   "value": gpairs, "unit": "Gpairs/s", "ms_per_step": hms, "scaling": "strong", "shards": world,
                   "exchange": ("none (one shard)" if world == 1 else "peer-memory stores over NVLink + arrival flags, fused with the shard merge; no collective call"
                                if use_peer else "one NCCL all-gather of 16 B per query per rank + merge kernel"),
                   "roofline": {"bound": "integer-pipe (POPC)", "achieved": gpairs, "peak": peak_gpairs, "unit": "Gpairs/s",
                                "frac": gpairs / peak_gpairs,
                                "note": "peak = measured POPC issue rate / 8 POPC per pair (algorithmic); the kernel's carry-save "
                                        "compression executes 4 POPC + 16 LOP3 per pair, so frac may exceed the naive bound",
                                "popc_per_s_measured": popc.value, "lop3_per_s_measured": lop3.value}}

    log("hamming done")
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        dom = max(("pyramid", "fast", "octree", "blur", "describe", "stereo"), key=lambda k: stages.get(k, 0.0))
        # algorithmic bytes of the dominant stage per launch: FAST reads every pyramid pixel once and writes 8 B per candidate;
        # the other stages' figures are listed in DESIGN.md
        alg_bytes = {"fast": LEVEL_PIXELS + 8 * 27400, "pyramid": LEVEL_PIXELS + (LEVEL_PIXELS - W * H), "blur": 2 * LEVEL_PIXELS,
                     "octree": 8 * 27000, "describe": 60 * 2006 + 2006 * (749 + 512), "stereo": 2006 * (64 * 2000 // 32)}[dom] * nimg
        achieved = alg_bytes / (stages[dom] * 1e-3) / 1e9 if stages.get(dom) else 0.0
        traffic = None      # dram__bytes_read + dram__bytes_write of the dominant kernel per launch, from the committed ncu capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_k_fast_cells_traffic.json")))
            if dom == "fast":
                traffic = int(tj["traffic_bytes_per_image"]) * nimg
        except Exception:
            pass
        pipeline_gbs = BYTES_PER_IMAGE * nimg / (ms_step * 1e-3) / 1e9
        cpu_baseline = None
        if not args.no_cpu_baseline:
            stereo_batch, kind, cpu_host = cpu_reference()
            cores = os.cpu_count() or 1
            sample_pairs = min(pairs, max(16, min(320, 12 * cores)))      # ~1 s of wall clock = 10-30 s of CPU work
            hl = host.numpy()
            stereo_batch(PARAMS, hl[:min(cores, sample_pairs)], hl[pairs:pairs + min(cores, sample_pairs)], BF, np.float32(mb), cores)   # warm
            t0 = time.perf_counter()
            stereo_batch(PARAMS, hl[:sample_pairs], hl[pairs:pairs + sample_pairs], BF, np.float32(mb), cores)
            cdt = time.perf_counter() - t0
            cpu_baseline = {"value": 2 * sample_pairs / cdt, "unit": "frames/s", "cores": cores, "kind": kind,
                            "sample": "%d of this run's stereo pairs, extraction L+R + ComputeStereoMatches, one pair per thread; %s"
                                      % (sample_pairs, cpu_host)}
        line = {
            "metric": "KITTI ORB extract frames/s @2000 kp", "value": fps, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "KITTI 1241x376 stereo pairs, nFeatures=2000, extract L+R + ComputeStereoMatches (BASELINE configs[1], batched)",
                       "pairs_per_step_per_gpu": pairs, "images_per_step_per_gpu": nimg, "mean_keypoints_per_image": mean_kp,
                       "l2": "inputs %.0f MB per step > 126 MB L2; pyramids %.0f MB" % (nimg * W * H / 1e6, nimg * 1.9),
                       "parallelism": "frame-sharded, %d rank(s), no data-path collective" % world,
                       "host_numa_node_rank0": numa_node},
            "clocks": clocks,
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "StereoBatchPipeline(depth=%d) over orbgpu_extract_batch_async + orbgpu_stereo_match_async + orbgpu_wait, pinned host buffers" % depth,
                    "single_pair_latency_ms": latency_ms,
                    "single_call_value": world * nimg * args.steps / dt_sync},
            "gpu_launches": launches,
            "stage_ms": stages,
            "roofline": {"bound": "hbm", "kernel": {"fast": "k_fast_cells", "pyramid": "k_resize_level", "octree": "k_octree", "blur": "k_blur_level",
                                                     "describe": "k_describe", "stereo": "k_stereo_match"}[dom],
                         "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic,
                         "peak_source": peak_src,
                         "pipeline": {"achieved": pipeline_gbs, "frac": pipeline_gbs / peak_gbs,
                                      "bytes_per_image": BYTES_PER_IMAGE, "note": "SURVEY 8d per-image bytes x images / whole step time"}},
            "cpu_baseline": cpu_baseline,
            "hamming": hamming,
        }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--pairs", type=int, default=320, help="stereo pairs per step per GPU (640 images = 299 MB > L2)")
    ap.add_argument("--ref-pairs", type=int, default=0)
    ap.add_argument("--no-hamming", action="store_true")
    ap.add_argument("--hamming-exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how the per-shard top-2 records travel: the library's peer-memory kernels (default) or one NCCL all-gather")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the informational single-pair latency loop (profiling runs)")
    args = ap.parse_args()
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if world == 1 and args.gpus > 1:
        # not under torchrun: re-launch one rank per GPU on this node
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.impl == "reference":
        return run_reference(args, rank, world)
    return run_own(args, rank, local_rank, world)


if __name__ == "__main__":
    sys.exit(main())
