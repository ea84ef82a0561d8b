#!/usr/bin/env python3
"""bench.py — ORB front-end throughput on B200 (BASELINE.json metric: ORB frames/s, 640x480, 1000 kp, 8 levels).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (hand-written sm_100a kernels via the C ABI)
  python bench.py --impl reference [...]                         the reference's CPU ORBextractor on the host cores
  torchrun --nproc-per-node N bench.py --gpus N ...              one rank per GPU, frames sharded by frame (weak scaling)

A step = one pass of the whole hot path (pyramid -> per-cell FAST -> quadtree -> orientation + blur + rBRIEF) over
one batch of synthetic frames. `value` = frames/s with the batch already resident in HBM; `e2e` = the same metric
through the public C-ABI call orbx_extract_batch with pinned HOST buffers (H2D + D2H inside the timed region).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

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

from orb_slam2_commit_b200 import synth  # noqa: E402

WORKLOADS = {
    # BASELINE.json metric is quoted on this one (configs[0] geometry, batched as in configs[2])
    "tum1": dict(cfg="tum1", batch=1024, distinct=32),
    "euroc": dict(cfg="euroc", batch=1024, distinct=32),
    "kitti": dict(cfg="kitti", batch=256, distinct=16),
    "4k": dict(cfg="4k", batch=32, distinct=4),
    # BASELINE configs[1]: stereo pairs, left + right extraction and Frame::ComputeStereoMatches
    "kitti_stereo": dict(cfg="kitti", batch=128, distinct=8, stereo=True),
    "euroc_stereo": dict(cfg="euroc", batch=256, distinct=8, stereo=True),
    # SURVEY §8(f) rows built around the extractor
    # stereo_euroc.cc:136-137: cv::remap rectification fused into pyramid level 0
    "euroc_rect": dict(cfg="euroc", batch=1024, distinct=32, rectify=True),
    # per-frame front-end of Tracking: extract -> UndistortKeyPoints -> ComputeBoW -> SearchByBoW / L1 score vs previous frame
    "tum1_frame": dict(cfg="tum1", batch=512, distinct=32, frame=True),
    # ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) on synthetic tracking scenes (1000 map points,
    # ~1100 current keypoints, 30 % near-duplicate map points): one CTA per frame pair
    "tum1_track": dict(cfg="tum1", batch=1024, distinct=8, track=True),
    # the other ORBmatcher functions, each batched one CTA per call: SearchByProjection(F, vpMapPoints) [SearchLocalPoints],
    # Fuse (search half), SearchByProjection(CurrentFrame, pKF, ..) [relocalisation], SearchForInitialization
    "tum1_matchers": dict(cfg="tum1", batch=1024, distinct=8, matchers=True),
}


# stdout carries exactly ONE JSON line. Libraries (NCCL's "NCCL version ..." banner, a chatty driver) write to fd 1 from
# C code, so main() points fd 1 at stderr for the whole run and the result line goes to the saved original stdout.
_RESULT_FD = None


# Host placement: a rank's pinned frame buffers should live on the NUMA node its GPU hangs off, otherwise every H2D copy
# crosses the socket interconnect and the end-to-end path of 4-8 ranks becomes host-memory bound. The rank binds itself to
# that node's cores before it allocates (first touch then lands locally) and restores the full mask before any CPU baseline.
_ORIG_AFFINITY = None
_NUMA_INFO = None


def bind_to_gpu_numa_node(torch, local):
    global _ORIG_AFFINITY, _NUMA_INFO
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            _NUMA_INFO = {"gpu_pci": bdf, "node": node, "bound": False}
            return
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        _ORIG_AFFINITY = os.sched_getaffinity(0)
        cpus &= _ORIG_AFFINITY
        if cpus:
            os.sched_setaffinity(0, cpus)
        _NUMA_INFO = {"gpu_pci": bdf, "node": node, "cpus": len(cpus), "bound": bool(cpus)}
    except Exception as e:                                  # placement is an optimisation, never a requirement
        _NUMA_INFO = {"bound": False, "error": str(e)[:80]}


def restore_affinity():
    if _ORIG_AFFINITY:
        try:
            os.sched_setaffinity(0, _ORIG_AFFINITY)
        except OSError:
            pass


def emit_json_line(line):
    if _NUMA_INFO is not None and isinstance(line.get("config"), dict):
        line["config"]["host_placement"] = _NUMA_INFO
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def level_geometry(c):
    """Pyramid payload sizes with the reference's own formula (ORBextractor.cc:1221-1225), float32 arithmetic."""
    sf = np.float32(1.0)
    sizes = []
    for l in range(c["nlevels"]):
        inv = np.float32(1.0) / sf
        sizes.append((int(np.rint(np.float32(c["width"]) * inv)), int(np.rint(np.float32(c["height"]) * inv))))
        sf = np.float32(np.float64(sf) * np.float64(np.float32(c["scale"])))
    return sizes


def algorithmic_bytes(c, nkp_mean):
    """per image B = P0 + 3A + N(961+60) for the reference's data flow; per stage of THIS design
    (DESIGN.md §kernels): pyramid P0 + A, FAST A, quadtree 8 B per candidate (in+out, ~ignored), describe 2A (dense blur:
    raw read + blurred write) + N*1021."""
    P0 = c["width"] * c["height"]
    A = sum(w * h for w, h in level_geometry(c))
    return dict(P0=P0, A=A, B_survey=P0 + 3 * A + nkp_mean * 1021, pyramid=P0 + A, fast=A, describe=2 * A + nkp_mean * 1021,
                quadtree=0)


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML polled from a thread every ~2 ms
    (nvidia-smi -lms needs longer to start than a short timed region lasts; it is only the fallback)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False
        self.thread = None
        self.h = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except (ValueError, IndexError):
                    pass
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.stop_flag = True
        self.thread.join(1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": float(self.max_mhz), "reasons": ["no samples"], "samples": 0}
        mhz = [m for m, _ in self.samples]
        bits = 0
        for _, r in self.samples:
            bits |= r
        reasons = sorted(n for b, n in self.REASONS.items() if bits & b)
        return {"sm_mhz": float(np.median(mhz)), "sm_max_mhz": float(self.max_mhz), "reasons": reasons, "samples": len(mhz)}


def make_frames(c, distinct, seed0=1):
    return np.stack([synth.synth_image(c["width"], c["height"], seed0 + i) for i in range(distinct)])


def cpu_reference_bench(c, frames, target_seconds, nthreads):
    """Times the reference's own CPU ORBextractor (oracle/_ref = verbatim ORBextractor.cc + cvshim) or, when that
    library is absent, the oracle port, frame-parallel over `nthreads` host threads. Returns (frames/s, kind, sample)."""
    restore_affinity()
    from oracle import binding as ob
    frames = np.ascontiguousarray(frames)
    n_img = len(frames)
    if ob.ref_available():
        R = ob.ref_lib()
        import ctypes as C
        tot = C.c_longlong(0)

        def run(iters):
            return R.orbref_bench(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"],
                                  frames.ctypes.data_as(ob.u8p), c["width"], c["height"], n_img, iters, nthreads, C.byref(tot))
        t1 = run(nthreads)                         # warm-up + calibration: one frame per thread
        iters = int(max(nthreads, min(100000, nthreads * max(1.0, target_seconds / max(t1, 1e-3)))))
        iters = (iters // nthreads) * nthreads
        t = run(iters)
        return iters / t, "reference", f"{iters} frames ({n_img} distinct) over {nthreads} threads, {t:.1f} s", tot.value / iters
    # port fallback: the oracle through ctypes (releases the GIL)
    exs = [ob.Extractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"]) for _ in range(nthreads)]
    counts = [0] * nthreads

    def work(t, iters):
        for i in range(t, iters, nthreads):
            k, _ = exs[t].extract(frames[i % n_img])
            counts[t] += len(k)

    def run(iters):
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t, iters)) for t in range(nthreads)]
        [x.start() for x in th]; [x.join() for x in th]
        return time.perf_counter() - t0
    t1 = run(nthreads)
    iters = int(max(nthreads, nthreads * max(1.0, target_seconds / max(t1, 1e-3))))
    counts[:] = [0] * nthreads
    t = run(iters)
    return iters / t, "port", f"{iters} frames ({n_img} distinct) over {nthreads} threads, {t:.1f} s", sum(counts) / iters


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    c = synth.CONFIGS[w["cfg"]]
    frames = make_frames(c, min(w["distinct"], 8))
    nthreads = os.cpu_count() or 1
    from oracle import binding as ob
    import ctypes as C
    per_step = 8 * nthreads                      # a bounded sample per step (~0.2 s on 16 threads), long enough that thread start-up does not weigh
    times = []
    kind = "reference" if ob.ref_available() else "port"
    if kind == "reference":
        R = ob.ref_lib(); tot = C.c_longlong(0)
        fr = np.ascontiguousarray(frames)
        for s in range(args.warmup + args.steps):
            t = R.orbref_bench(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"], fr.ctypes.data_as(ob.u8p),
                               c["width"], c["height"], len(fr), per_step, nthreads, C.byref(tot))
            if s >= args.warmup:
                times.append(t)
    else:
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            fps, _, _, _ = cpu_reference_bench(c, frames, 0.0, nthreads)
            if s >= args.warmup:
                times.append(per_step / fps)
    total = sum(times)
    value = per_step * len(times) / total
    line = {"impl": "reference", "metric": "orb_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{w['cfg']}: {c['width']}x{c['height']}, nFeatures={c['nfeatures']}, scale {c['scale']}, {c['nlevels']} levels, FAST {c['ini_th']}/{c['min_th']}",
                       "frames_per_step": per_step, "host_threads": nthreads},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": nthreads, "kind": kind,
                             "sample": f"{per_step} frames per step, {len(times)} steps"},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_json_line(line)


def bench_hamming(torch, api_lib, dev, peaks, sm_mhz):
    """BASELINE config 4: 2048 x 1,000,000 x 256-bit brute-force top-2, inputs resident."""
    nq, nt = 2048, 1_000_000
    g = torch.Generator(device=dev); g.manual_seed(42)
    train = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, device=dev, generator=g)
    query = train[torch.randint(0, nt, (nq,), device=dev, generator=g)].clone()
    query[:, :4] ^= 0x5a
    packed = torch.empty(nq, dtype=torch.int64, device=dev)
    out = torch.empty(3, nq, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def run():
        api_lib.orbx_hamming_init_device(packed.data_ptr(), nq, st)
        api_lib.orbx_hamming_top2_device(query.data_ptr(), nq, train.data_ptr(), nt, 0, packed.data_ptr(), st)
        api_lib.orbx_hamming_merge_device(packed.data_ptr(), 1, nq, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), st)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    alg_bytes = 32 * nt + 32 * nq + 8 * nq
    clk = (sm_mhz or 1965.0) * 1e6
    # what binds (ncu, profiles/r02_ncu_hamming.txt): the integer-logic pipe. Per (query, train row) pair the carry-save tree
    # costs 8 XOR + 8 three-input logic ops + 1 key + 3 packed-key min/max = 20 ALU-pipe instructions (64 lanes per clock per
    # SM) and 4 POPC (16 lanes per clock per SM); eight plain POPC per pair, the round-1 form, bound the kernel at 3.5 ms.
    alu_ops, popc_ops = 20.0 * nq * nt, 4.0 * nq * nt
    return {"workload": "2048 queries x 1,000,000 train rows, 256-bit, top-2", "ms": ms, "matches_per_s": nq / (ms * 1e-3),
            "pair_distances_per_s": nq * nt / (ms * 1e-3), "algorithmic_GBps": alg_bytes / (ms * 1e-3) / 1e9,
            "hbm_frac": alg_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "bound": "alu pipe", "alu_pipe_frac": alu_ops / (ms * 1e-3) / (148 * 64 * clk),
            "popc_pipe_frac": popc_ops / (ms * 1e-3) / (148 * 16 * clk),
            "pipe_model": "148 SMs x 64 ALU lanes (16 POPC lanes) per clock x measured SM clock; 20 ALU + 4 POPC instructions per pair (Harley-Seal carry-save tree)",
            "round1_plain_popc_ms": 3.877}


def cpu_thread_bench(make_worker, target_seconds, nthreads):
    """Generic CPU baseline: make_worker(tid) -> f(i) doing one unit of work through ctypes (GIL released)."""
    restore_affinity()
    workers = [make_worker(t) for t in range(nthreads)]

    def run(iters):
        def work(t):
            for i in range(t, iters, nthreads):
                workers[t](i)
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
        [x.start() for x in th]; [x.join() for x in th]
        return time.perf_counter() - t0
    t1 = run(nthreads)
    iters = int(max(nthreads, nthreads * max(1.0, target_seconds / max(t1, 1e-3))))
    tt = run(iters)
    return iters / tt, f"{iters} units over {nthreads} threads, {tt:.1f} s"


def run_frame(args, torch, dist, rank, world, local, dev):
    """Frame front-end behind the extractor, chained on the device with no host round trip:
    extract -> Frame::UndistortKeyPoints -> Frame::ComputeBoW (vocabulary k=10, L=6, levelsup 4) -> for every frame
    ORBmatcher::SearchByBoW against the previous frame of the batch and the L1 BoW score against it."""
    from orb_slam2_commit_b200 import ORBextractor, ORBVocabulary, api
    w = WORKLOADS[args.workload]; c = synth.CONFIGS[w["cfg"]]
    B = args.batch or w["batch"]; W, H = c["width"], c["height"]
    frames = make_frames(c, w["distinct"], seed0=1 + 1000 * rank)
    host_batch = torch.empty((B, H, W), dtype=torch.uint8, pin_memory=True); hb = host_batch.numpy()
    for i in range(B):
        hb[i] = frames[i % len(frames)]
    d_imgs = host_batch.to(dev)
    voc = synth.synth_vocabulary(10, 6, 7)
    V = ORBVocabulary(10, 6, *voc, device=local)
    cfgargs = (c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    ex = ORBextractor(*cfgargs, device=local)
    cap = ex.reserve(W, H, B)
    mk = lambda *shape, dt=torch.uint8: torch.zeros(shape, dtype=dt, device=dev)
    d_kps, d_un, d_desc, d_nkp = mk(B, cap, 28), mk(B, cap, 28), mk(B, cap, 32), mk(B, dt=torch.int32)
    # every frame is matched against the previous occurrence of the same scene in the stream (the batch cycles through
    # `distinct` synthetic scenes), so SearchByBoW / the score see a realistic overlap instead of unrelated images
    kf = torch.arange(0, B, dtype=torch.int32, device=dev); ff = torch.roll(kf, -len(frames))
    d_match, d_nm, d_score = mk(B, cap, dt=torch.int32), mk(B, dt=torch.int32), mk(B, dt=torch.float64)
    K4 = synth.TUM1_K4; D5 = synth.TUM1_DIST
    import ctypes as C
    f32p = C.POINTER(C.c_float)
    L = api.lib()
    ts = torch.cuda.Stream(device=dev); torch.cuda.set_stream(ts); st = ts.cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]

    def step(timed=False):
        if timed: ev[0].record()
        ex.extract_device(d_imgs.data_ptr(), B, W, H, W, W * H, d_kps.data_ptr(), cap, d_nkp.data_ptr(), d_desc.data_ptr(), st)
        if timed: ev[1].record()
        api._ck(L.orbx_undistort_keypoints_device(d_kps.data_ptr(), B * cap, K4.ctypes.data_as(f32p), D5.ctypes.data_as(f32p), 5,
                                                  d_un.data_ptr(), st))
        if timed: ev[2].record()
        V.transform_device(d_desc.data_ptr(), d_nkp.data_ptr(), B, cap, 4, st)
        if timed: ev[3].record()
        V.search_by_bow_device(B, kf.data_ptr(), ff.data_ptr(), d_un.data_ptr(), d_desc.data_ptr(), None, 0.7, True,
                               d_match.data_ptr(), d_nm.data_ptr(), st)
        V.score_device(kf.data_ptr(), ff.data_ptr(), B, d_score.data_ptr(), st)
        if timed: ev[4].record()

    def barrier():
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup): step()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps): step()
    e1.record(); torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1); clocks = sampler.stop()
    step(timed=True); torch.cuda.synchronize()
    st_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * args.steps / (ms_total * 1e-3)
    outs_src = (d_un, d_desc, d_nkp, d_match, d_nm, d_score)
    outs = [torch.empty_like(x, device="cpu").pin_memory() for x in outs_src]

    def e2e_step():
        d_imgs.copy_(host_batch, non_blocking=True)
        step()
        for o, x in zip(outs, outs_src): o.copy_(x, non_blocking=True)
        torch.cuda.synchronize()
    for _ in range(2): e2e_step()
    barrier()
    n_e2e = max(3, min(args.steps, 10)); t0 = time.perf_counter()
    for _ in range(n_e2e): e2e_step()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_v = world * B * n_e2e / float(t.item())
    if rank != 0: return
    nkp = outs[2].numpy()
    line = {"metric": "frontend_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": f"{w['cfg']} frame front-end: {W}x{H}, nFeatures={c['nfeatures']}: extract + UndistortKeyPoints + ComputeBoW "
                                   f"(synthetic vocabulary k=10 L=6, {V.nwords} words) + SearchByBoW and L1 score against the previous frame",
                       "frames_per_step_per_gpu": B, "distinct_frames": len(frames)},
            "clocks": clocks, "gpu_launches": (c["nlevels"] + 4 + 1 + 2 + 3 + 1) * args.steps,
            "e2e": {"value": e2e_v, "unit": "frames/s", "h2d_bytes_per_step": B * W * H,
                    "d2h_bytes_per_step": int(sum(o.numel() * o.element_size() for o in outs)), "steps": n_e2e,
                    "api": "orbx_extract_device + orbx_undistort_keypoints_device + orbx_bow_transform_device + orbx_search_by_bow_device + orbx_bow_score_device, pinned host buffers"},
            "stages": {"extract_ms": st_ms[0], "undistort_ms": st_ms[1], "bow_transform_ms": st_ms[2], "search_by_bow_and_score_ms": st_ms[3]},
            "pipeline": {"keypoints_per_frame": float(nkp.mean()), "bow_matches_per_frame": float(outs[4].numpy().mean()),
                         "mean_l1_score_vs_previous": float(outs[5].numpy().mean())}}
    if world == 1 and not args.no_cpu_baseline:
        from oracle import binding as ob
        nthreads = os.cpu_count() or 1
        Vo = ob.Vocabulary(10, 6, *voc)
        def make_worker(tid):
            e = ob.Extractor(*cfgargs)
            state = {}
            def f(i):
                k, d = e.extract(frames[i % len(frames)])
                ob.undistort_points(np.stack([k["x"], k["y"]], 1), K4, D5)
                tcur = Vo.transform(d, 4)
                if "t" in state:
                    ob.search_by_bow(state["t"], tcur, state["d"], state["k"]["angle"], np.ones(len(state["k"]), np.uint8), d, k["angle"], 0.7, True)
                    ob.bow_score_l1(state["t"]["bow_id"], state["t"]["bow_val"], tcur["bow_id"], tcur["bow_val"])
                state.update(t=tcur, d=d, k=k)
            return f
        v, sample = cpu_thread_bench(make_worker, args.cpu_seconds, nthreads)
        line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": nthreads, "kind": "port", "sample": sample}
    emit_json_line(line)


def run_track(args, torch, dist, rank, world, local, dev):
    """Frame-to-frame SearchByProjection (ORBmatcher.cc:1489-1646) for a batch of frame pairs resident in HBM."""
    import ctypes as C
    from orb_slam2_commit_b200 import api, search_by_projection_frame
    w = WORKLOADS[args.workload]
    P = args.batch or w["batch"]
    scenes = [synth.synth_tracking_scene(11 + i + 100 * rank) for i in range(w["distinct"])]
    L = api.lib()
    keep = []                                   # device tensors that must outlive the pair structs

    def dev_t(a):
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev); keep.append(t); return t
    dsc = []
    for sc in scenes:
        d = {k: dev_t(sc[k]) for k in ("cur_kps", "cur_desc", "cur_u_right", "cur_occupied", "last_kps", "last_xyz", "last_desc", "last_flags")}
        dsc.append(d)
    pairs = (api.OrbxProjectionPair * P)()
    d_match = torch.zeros((P, max(len(s_["cur_kps"]) for s_ in scenes)), dtype=torch.int32, device=dev)
    d_nm = torch.zeros(P, dtype=torch.int32, device=dev)
    for i in range(P):
        sc, d = scenes[i % len(scenes)], dsc[i % len(scenes)]
        p = pairs[i]
        p.cur_keypoints = d["cur_kps"].data_ptr(); p.cur_descriptors = d["cur_desc"].data_ptr()
        p.cur_u_right = d["cur_u_right"].data_ptr(); p.cur_occupied = d["cur_occupied"].data_ptr(); p.n_cur = len(sc["cur_kps"])
        p.last_keypoints = d["last_kps"].data_ptr(); p.last_xyz = d["last_xyz"].data_ptr()
        p.last_descriptors = d["last_desc"].data_ptr(); p.last_flags = d["last_flags"].data_ptr(); p.n_last = len(sc["last_kps"])
        p.Tcw = (C.c_float * 12)(*sc["Tcw12"].tolist()); p.mode = i % 3
        p.match = d_match[i].data_ptr(); p.nmatches = d_nm[i:].data_ptr()
    cam = scenes[0]["cam9"]; sf = scenes[0]["scale_factors"]
    f32p = C.POINTER(C.c_float)
    ts = torch.cuda.Stream(device=dev); torch.cuda.set_stream(ts); st = ts.cuda_stream

    def step():
        api._ck(L.orbx_search_by_projection_device(pairs, P, cam.ctypes.data_as(f32p), sf.ctypes.data_as(f32p), len(sf), 7.0, 1, local, st))

    def barrier():
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup): step()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps): step()
    e1.record(); torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1); clocks = sampler.stop()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * P * args.steps / (ms_total * 1e-3)
    # e2e: the synchronous single-pair host call (host buffers in, assignment out)
    n_e2e = 200; t0 = time.perf_counter()
    for i in range(n_e2e):
        search_by_projection_frame(**scenes[i % len(scenes)], th=7.0, mode=i % 3, device=local)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_v = world * n_e2e / float(t.item())
    if rank != 0: return
    sc0 = scenes[0]
    line = {"metric": "projection_search_pairs_per_s", "value": value, "unit": "frame pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"SearchByProjection(CurrentFrame, LastFrame, th=7): {len(sc0['last_kps'])} last-frame map points x "
                                   f"{len(sc0['cur_kps'])} current keypoints per pair, rotation check on, modes cycled",
                       "pairs_per_step_per_gpu": P, "distinct_scenes": len(scenes)},
            "clocks": clocks, "gpu_launches": args.steps,
            "e2e": {"value": e2e_v, "unit": "frame pairs/s", "h2d_bytes_per_step": int(sum(np.asarray(v).nbytes for v in sc0.values() if v is not None)),
                    "d2h_bytes_per_step": 4 * len(sc0["cur_kps"]) + 4, "steps": n_e2e, "api": "orbx_search_by_projection (one pair per call, synchronous)"},
            "pipeline": {"matches_per_pair": float(d_nm.float().mean().item())}}
    if world == 1 and not args.no_cpu_baseline:
        from oracle import binding as ob
        nthreads = os.cpu_count() or 1
        def make_worker(tid):
            return lambda i: ob.search_by_projection_frame(**scenes[i % len(scenes)], th=7.0, mode=i % 3)
        v, sample = cpu_thread_bench(make_worker, min(args.cpu_seconds, 6.0), nthreads)
        line["cpu_baseline"] = {"value": v, "unit": "frame pairs/s", "cores": nthreads, "kind": "port", "sample": sample}
    emit_json_line(line)


def run_matchers(args, torch, dist, rank, world, local, dev):
    """The window matchers of ORBmatcher beyond the frame-to-frame one, device-resident and batched (one CTA per call):
    headline = SearchByProjection(F, vpMapPoints, th) as Tracking::SearchLocalPoints calls it; the others in `others`."""
    import ctypes as C
    from orb_slam2_commit_b200 import api, search_local_points
    w = WORKLOADS[args.workload]
    P = args.batch or w["batch"]
    L = api.lib()
    keep = []
    f32p = C.POINTER(C.c_float)

    def dev_t(a):
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev); keep.append(t); return t.data_ptr()

    def fp(a):
        a = np.ascontiguousarray(a, np.float32); keep.append(a); return a.ctypes.data_as(f32p)
    ts = torch.cuda.Stream(device=dev); torch.cuda.set_stream(ts); st = ts.cuda_stream
    nd = w["distinct"]
    # ---- SearchLocalPoints
    lp = [synth.synth_local_points_scene(21 + i + 100 * rank) for i in range(nd)]
    lpd = [{k: (None if v is None else dev_t(v)) for k, v in sc.items() if k not in ("bounds4", "scale_factors")} for sc in lp]
    frames = (api.OrbxLocalPointsFrame * P)()
    d_match = torch.zeros((P, max(len(s_["kps"]) for s_ in lp)), dtype=torch.int32, device=dev)
    d_nm = torch.zeros(P, dtype=torch.int32, device=dev)
    for i in range(P):
        sc, d = lp[i % nd], lpd[i % nd]; f = frames[i]
        f.keypoints = d["kps"]; f.descriptors = d["desc"]; f.u_right = d["u_right"]; f.occupied = d["occupied"]; f.n = len(sc["kps"])
        f.queries = d["queries"]; f.query_descriptors = d["query_desc"]; f.query_flags = d["query_flags"]; f.nq = len(sc["queries"])
        f.match = d_match[i].data_ptr(); f.nmatches = d_nm[i:].data_ptr()
    b4, sf = fp(lp[0]["bounds4"]), fp(lp[0]["scale_factors"])

    def step_local():
        api._ck(L.orbx_search_local_points_device(frames, P, b4, sf, 8, 1.0, 0.8, local, st))
    # ---- Fuse (search half) and the relocalisation matcher share one scene family
    fs = [synth.synth_kf_projection_scene(31 + i + 100 * rank) for i in range(nd)]
    fsd = [{k: dev_t(v) for k, v in sc.items() if k in ("kps", "desc", "occupied", "pt_xyz", "pt_normal", "pt_dist", "pt_desc", "pt_flags", "pt_angle")} for sc in fs]
    fj = (api.OrbxFuseJob * P)(); pj = (api.OrbxProjectionJob * P)()
    npt = max(len(s_["pt_flags"]) for s_ in fs)
    d_bi = torch.zeros((P, npt), dtype=torch.int32, device=dev); d_bd = torch.zeros((P, npt), dtype=torch.int32, device=dev)
    d_nf = torch.zeros(P, dtype=torch.int32, device=dev)
    d_m2 = torch.zeros((P, max(len(s_["kps"]) for s_ in fs)), dtype=torch.int32, device=dev); d_nm2 = torch.zeros(P, dtype=torch.int32, device=dev)
    for i in range(P):
        sc, d = fs[i % nd], fsd[i % nd]
        for j in (fj[i], pj[i]):
            j.keypoints = d["kps"]; j.descriptors = d["desc"]; j.n = len(sc["kps"])
            j.Tcw = (C.c_float * 12)(*sc["Tcw12"].tolist()); j.Ow = (C.c_float * 3)(*sc["Ow3"].tolist())
            j.pt_xyz = d["pt_xyz"]; j.pt_normal = d["pt_normal"]; j.pt_dist = d["pt_dist"]; j.pt_descriptors = d["pt_desc"]
            j.pt_flags = d["pt_flags"]; j.npts = len(sc["pt_flags"]); j.mode = 0
        fj[i].u_right = None; fj[i].th = 3.0
        fj[i].best_idx = d_bi[i].data_ptr(); fj[i].best_dist = d_bd[i].data_ptr(); fj[i].nfused = d_nf[i:].data_ptr()
        pj[i].occupied = d["occupied"]; pj[i].pt_angle = d["pt_angle"]; pj[i].th = 10.0; pj[i].max_dist = 100
        pj[i].match = d_m2[i].data_ptr(); pj[i].nmatches = d_nm2[i:].data_ptr()
    cam9, sff, is2 = fp(fs[0]["cam9"]), fp(fs[0]["scale_factors"]), fp(1.0 / (fs[0]["scale_factors"] ** 2))
    lsf = fs[0]["log_scale_factor"]

    def step_fuse():
        api._ck(L.orbx_fuse_search_device(fj, P, cam9, sff, is2, 8, lsf, local, st))

    def step_reloc():
        api._ck(L.orbx_search_by_projection_kf_device(pj, P, cam9, sff, 8, lsf, 1, local, st))
    # ---- SearchForInitialization
    ini = [synth.synth_initialization_scene(41 + i + 100 * rank) for i in range(nd)]
    inid = [{k: dev_t(v) for k, v in sc.items() if k != "bounds4"} for sc in ini]
    ip = (api.OrbxInitPair * P)()
    d_m3 = torch.zeros((P, max(len(s_["kps1"]) for s_ in ini)), dtype=torch.int32, device=dev); d_nm3 = torch.zeros(P, dtype=torch.int32, device=dev)
    d_prev = torch.zeros((P, max(len(s_["kps1"]) for s_ in ini), 2), dtype=torch.float32, device=dev)
    for i in range(P):
        sc, d = ini[i % nd], inid[i % nd]; q = ip[i]
        q.keypoints1 = d["kps1"]; q.descriptors1 = d["desc1"]; q.n1 = len(sc["kps1"]); q.keypoints2 = d["kps2"]; q.descriptors2 = d["desc2"]
        q.n2 = len(sc["kps2"]); q.prev_matched = d["prev_matched"]; q.prev_matched_out = d_prev[i].data_ptr(); q.window_size = 100
        q.match12 = d_m3[i].data_ptr(); q.nmatches = d_nm3[i:].data_ptr()
    ib4 = fp(ini[0]["bounds4"])

    def step_init():
        api._ck(L.orbx_search_for_initialization_device(ip, P, ib4, 0.9, 1, local, st))

    # ---- SearchForTriangulation: nd keyframe pairs transformed once (ComputeBoW), then P pair matches per launch
    from orb_slam2_commit_b200 import ORBVocabulary
    voc = synth.synth_vocabulary(10, 4, 5)
    V = ORBVocabulary(10, 4, *voc, device=local)
    tri = [synth.synth_triangulation_scene(voc, 51 + i + 100 * rank) for i in range(nd)]
    tcap = max(max(len(t["kps1"]), len(t["kps2"])) for t in tri)
    t_desc = np.zeros((2 * nd, tcap, 32), np.uint8); t_kps = np.zeros((2 * nd, tcap), api.KP_DTYPE); t_cnt = np.zeros(2 * nd, np.int32)
    t_mp = np.zeros((2 * nd, tcap), np.uint8); t_geom = np.zeros((P, 28), np.float32)
    for i, t in enumerate(tri):
        for j, tag in enumerate(("1", "2")):
            n_ = len(t["kps" + tag]); t_cnt[2 * i + j] = n_
            t_desc[2 * i + j, :n_] = t["desc" + tag]; t_kps[2 * i + j, :n_] = t["kps" + tag]; t_mp[2 * i + j, :n_] = t["has_mp" + tag]
    for i in range(P):
        t_geom[i] = tri[i % nd]["geom28"]
    d_tdesc = dev_t(t_desc); d_tkps = dev_t(t_kps); d_tmp = dev_t(t_mp); d_tgeom = dev_t(t_geom)
    d_tcnt = torch.from_numpy(t_cnt).to(dev)
    d_f1 = torch.tensor([2 * (i % nd) for i in range(P)], dtype=torch.int32, device=dev)
    d_f2 = torch.tensor([2 * (i % nd) + 1 for i in range(P)], dtype=torch.int32, device=dev)
    d_m4 = torch.zeros((P, tcap), dtype=torch.int32, device=dev); d_nm4 = torch.zeros(P, dtype=torch.int32, device=dev)
    V.transform_device(d_tdesc, d_tcnt.data_ptr(), 2 * nd, tcap, 2, st)
    tsf, ts2 = fp(tri[0]["scale_factors"]), fp(tri[0]["level_sigma2"])

    def step_tri():
        api._ck(L.orbx_search_for_triangulation_device(V._h, P, d_f1.data_ptr(), d_f2.data_ptr(), d_tkps, d_tdesc, d_tmp, None, d_tgeom,
                                                       tsf, ts2, 8, 0, 1, d_m4.data_ptr(), d_nm4.data_ptr(), st))
    # ---- SearchBySim3: one keyframe pair per call (two projection jobs + the mutual check in two launches)
    s3 = synth.synth_sim3_scene(61 + 100 * rank)
    ks = []
    for k in s3[:2]:
        K = api.OrbxSim3KeyFrame()
        K.keypoints = dev_t(k["kps"]); K.descriptors = dev_t(k["desc"]); K.n = len(k["kps"]); K.mp_xyz = dev_t(k["mp_xyz"])
        K.mp_dist = dev_t(k["mp_dist"]); K.mp_descriptors = dev_t(k["mp_desc"]); K.mp_flags = dev_t(k["mp_flags"])
        K.Tcw = (C.c_float * 12)(*k["Tcw12"].tolist()); ks.append(K)
    d_m5 = torch.zeros(ks[0].n, dtype=torch.int32, device=dev); d_nf5 = torch.zeros(1, dtype=torch.int32, device=dev)
    d_scr = torch.zeros(ks[0].n + ks[1].n + 2, dtype=torch.int32, device=dev)
    s12, s21, scam, ssf = fp(s3[2]), fp(s3[3]), fp(s3[4]), fp(s3[5])
    n_sim3 = 64

    def step_sim3():
        for _ in range(n_sim3):
            api._ck(L.orbx_search_by_sim3_device(C.byref(ks[0]), C.byref(ks[1]), s12, s21, scam, ssf, 8, s3[6], 7.5, d_m5.data_ptr(),
                                                 d_nf5.data_ptr(), d_scr.data_ptr(), local, st))

    def barrier():
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()

    def timed(step, steps):
        for _ in range(args.warmup): step()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps): step()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    sampler = ClockSampler(local); sampler.start()
    ms_total = timed(step_local, args.steps)
    clocks = sampler.stop()
    value = world * P * args.steps / (ms_total * 1e-3)
    others = {}
    for name, step in (("fuse_search", step_fuse), ("search_by_projection_kf_relocalisation", step_reloc), ("search_for_initialization", step_init)):
        ms = timed(step, max(5, args.steps // 5))
        others[name] = {"calls_per_s": world * P * max(5, args.steps // 5) / (ms * 1e-3), "ms_per_launch": ms / max(5, args.steps // 5), "calls_per_launch": P}
    ms = timed(step_tri, max(5, args.steps // 5))
    others["search_for_triangulation"] = {"calls_per_s": world * P * max(5, args.steps // 5) / (ms * 1e-3), "ms_per_launch": ms / max(5, args.steps // 5),
                                          "calls_per_launch": P, "matches_per_call": float(d_nm4.float().mean().item())}
    ms = timed(step_sim3, 5)
    others["search_by_sim3"] = {"calls_per_s": world * n_sim3 * 5 / (ms * 1e-3), "ms_per_call": ms / (5 * n_sim3), "found_per_call": int(d_nf5.item()),
                                "note": "one keyframe pair per call, stream-ordered, no synchronisation between calls"}
    others["fuse_search"]["fused_per_call"] = float(d_nf.float().mean().item())
    others["search_by_projection_kf_relocalisation"]["matches_per_call"] = float(d_nm2.float().mean().item())
    others["search_for_initialization"]["matches_per_call"] = float(d_nm3.float().mean().item())
    n_e2e = 200; t0 = time.perf_counter()
    for i in range(n_e2e):
        search_local_points(**lp[i % nd], th=1.0, nnratio=0.8, device=local)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_v = world * n_e2e / float(t.item())
    # ---- the same call against a device-resident Frame handle (orbx_frame_*): keypoints / descriptors / mvuRight stay in HBM,
    # only the projected map points go up (one packed copy from pinned staging) and the assignments come down. One handle per
    # host thread (own stream): (a) one thread, call after call = the latency a tracking thread sees; (b) all host threads at
    # once, each on its own handle = what the machine sustains — the form the 16-thread CPU baseline is measured in.
    from orb_slam2_commit_b200 import Frame
    nthreads_h = os.cpu_count() or 1
    handles = []
    for t_ in range(nthreads_h):
        sc = lp[t_ % nd]
        fr = Frame(len(sc["kps"]), len(sc["queries"]), device=local)
        fr.from_device(lpd[t_ % nd]["kps"], lpd[t_ % nd]["desc"], len(sc["kps"]), stream=st)
        fr.set_stereo(sc["u_right"])
        handles.append((fr, sc, np.empty(len(sc["kps"]), np.int32)))
    torch.cuda.synchronize()

    def frame_call(h):
        fr, sc, out = h
        return fr.search_local_points(sc["queries"], sc["query_desc"], sc["query_flags"], sc["occupied"], sc["bounds4"], sc["scale_factors"], 1.0, 0.8, out=out)
    want_nm = [search_local_points(**lp[i], th=1.0, nnratio=0.8, device=local)[0] for i in range(nd)]
    for t_, h in enumerate(handles):
        assert frame_call(h)[0] == want_nm[t_ % nd], "frame-handle result differs from the one-shot host call"
    t0 = time.perf_counter()
    for i in range(400):
        frame_call(handles[0])
    frame_lat_ms = (time.perf_counter() - t0) / 400 * 1e3
    per_thread = 300

    def worker(h):
        for _ in range(per_thread):
            frame_call(h)
    barrier()
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(h,)) for h in handles]
    [x.start() for x in th]; [x.join() for x in th]
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    frame_mt_v = world * nthreads_h * per_thread / float(t.item())
    del handles
    if rank != 0: return
    sc0 = lp[0]
    line = {"metric": "search_local_points_calls_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"SearchByProjection(F, vpMapPoints, th=1): {len(sc0['queries'])} local map points x {len(sc0['kps'])} keypoints "
                                   f"per frame (Tracking::SearchLocalPoints), 30 % near-duplicate points, stereo gate on",
                       "frames_per_step_per_gpu": P, "distinct_scenes": nd},
            "clocks": clocks, "gpu_launches": args.steps,
            "e2e": {"value": e2e_v, "unit": "frames/s", "h2d_bytes_per_step": int(sum(np.asarray(v).nbytes for v in sc0.values() if v is not None)),
                    "d2h_bytes_per_step": 4 * len(sc0["kps"]) + 4, "steps": n_e2e, "api": "orbx_search_local_points (one frame per call, synchronous)",
                    "frame_handle": {"api": "orbx_frame_search_local_points (device-resident Frame, one frame per call, synchronous)",
                                     "h2d_bytes_per_call": int(sc0["queries"].nbytes + np.asarray(sc0["query_desc"]).nbytes + len(sc0["query_flags"]) + len(sc0["kps"]) + 256),
                                     "one_host_thread": {"frames_per_s": 1e3 / frame_lat_ms, "ms_per_call": frame_lat_ms},
                                     "all_host_threads": {"frames_per_s": frame_mt_v, "threads": nthreads_h, "handles": nthreads_h,
                                                          "note": "every thread calls on its own handle / stream; the kernels of different handles overlap on the GPU"}}},
            "pipeline": {"matches_per_frame": float(d_nm.float().mean().item())}, "others": others}
    if world == 1:
        # the same calls from NATIVE host threads (tools/native/bench_frame_threads.cu, built here with nvcc): the Python threads
        # above share one interpreter lock and spend ~25 us per call in it, which caps that leg near 40 k calls/s
        try:
            import subprocess, tempfile
            tmp = tempfile.mkdtemp(prefix="orbx_native_")
            exe, scene = os.path.join(tmp, "bench_frame_threads"), os.path.join(tmp, "lp_scene.bin")
            csrc = os.path.join(ROOT, "orb_slam2_commit_b200", "csrc")
            subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, os.path.join(ROOT, "tools", "native", "bench_frame_threads.cu"),
                            "-I" + os.path.join(ROOT, "include"), "-L" + csrc, "-lorbx", "-Xlinker", "-rpath", "-Xlinker", csrc],
                           check=True, capture_output=True, timeout=300)
            subprocess.run([sys.executable, os.path.join(ROOT, "tools", "native", "dump_scene.py"), scene], check=True, capture_output=True, timeout=300)
            out_ = subprocess.run([exe, scene, str(nthreads_h), "2000"], check=True, capture_output=True, text=True, timeout=300).stdout
            nat = json.loads(out_.strip().splitlines()[-1])
            line["e2e"]["frame_handle"]["native_host_threads"] = {"frames_per_s": nat["all_threads_calls_per_s"], "threads": nat["threads"],
                                                                   "one_thread_frames_per_s": nat["one_thread_calls_per_s"],
                                                                   "how": "tools/native/bench_frame_threads.cu: std::thread callers, one handle each, no interpreter between the calls"}
        except Exception as e:  # the tool is optional: never lose the line
            line["e2e"]["frame_handle"]["native_host_threads"] = {"error": str(e)[:200]}
    if world == 1 and not args.no_cpu_baseline:
        from oracle import binding as ob
        nthreads = os.cpu_count() or 1
        secs = min(args.cpu_seconds, 4.0)
        v, sample = cpu_thread_bench(lambda tid: (lambda i: ob.search_local_points(**lp[i % nd], th=1.0, nnratio=0.8)), secs, nthreads)
        line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": nthreads, "kind": "port", "sample": sample}
        v, _ = cpu_thread_bench(lambda tid: (lambda i: ob.fuse_search(**{k: x for k, x in fs[i % nd].items() if k not in ("occupied", "pt_angle")}, u_right=None, inv_level_sigma2=1.0 / (fs[0]["scale_factors"] ** 2), th=3.0, mode=0)), 2.0, nthreads)
        others["fuse_search"]["cpu_calls_per_s"] = v
        v, _ = cpu_thread_bench(lambda tid: (lambda i: ob.search_by_projection_kf(**fs[i % nd], th=10.0, max_dist=100, mode=0)), 2.0, nthreads)
        others["search_by_projection_kf_relocalisation"]["cpu_calls_per_s"] = v
        v, _ = cpu_thread_bench(lambda tid: (lambda i: ob.search_for_initialization(**ini[i % nd])), 2.0, nthreads)
        others["search_for_initialization"]["cpu_calls_per_s"] = v
    emit_json_line(line)


def h2d_ceiling_gbs(torch, dist, host_batch, dev, world, reps=8, d2h_bytes=0):
    """What the box can feed: every rank copies its own pinned batch to its GPU at the same time, no kernels running
    (cudaMemcpyAsync on one stream, `reps` copies back to back); aggregate GB/s = world x bytes x reps / max-over-ranks time.
    With d2h_bytes > 0 a second stream copies that many bytes per rep back into pinned memory at the same time (the results of
    a step flow back while the next frames go up): returns (h2d-only GB/s, duplex total GB/s)."""
    dst = torch.empty_like(host_batch, device=dev)
    st = torch.cuda.Stream(device=dev); st2 = torch.cuda.Stream(device=dev)
    back_d = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, device=dev)
    back_h = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, pin_memory=True)

    def run(duplex):
        with torch.cuda.stream(st):
            dst.copy_(host_batch, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(st):
                dst.copy_(host_batch, non_blocking=True)
            if duplex:
                with torch.cuda.stream(st2):
                    back_h.copy_(back_d, non_blocking=True)
        st.synchronize(); st2.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    up = world * host_batch.numel() * reps / run(False) / 1e9
    both = world * (host_batch.numel() + d2h_bytes) * reps / run(True) / 1e9 if d2h_bytes else None
    return up, both


def stereo_measure(wname, P, steps, warmup, torch, dist, rank, world, local, dev, distinct=None):
    """Stereo pairs/s: two extractor instances (as the reference, Tracking.cc:120-123) + the device-resident matcher
    (Frame.cc:80-84, 547-788); pairs shard by frame over the ranks, no collective. Returns the bench line (without a CPU
    baseline) on rank 0 and None elsewhere; every rank takes part in the barriers / reductions."""
    from orb_slam2_commit_b200 import ORBextractor, api, stereo_match_device
    w = WORKLOADS[wname]; c = synth.CONFIGS[w["cfg"]]
    W, H = c["width"], c["height"]
    pairs = [synth.synth_stereo_pair(W, H, 2 + i + 1000 * rank) for i in range(distinct or w["distinct"])]
    args = argparse.Namespace(steps=steps, warmup=warmup)
    hL = torch.empty((P, H, W), dtype=torch.uint8, pin_memory=True); hR = torch.empty((P, H, W), dtype=torch.uint8, pin_memory=True)
    for i in range(P):
        hL.numpy()[i] = pairs[i % len(pairs)][0]; hR.numpy()[i] = pairs[i % len(pairs)][1]
    dL, dR = hL.to(dev), hR.to(dev)
    cfgargs = (c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    exL, exR = ORBextractor(*cfgargs, device=local), ORBextractor(*cfgargs, device=local)
    cap = exL.reserve(W, H, P); exR.reserve(W, H, P)
    mk = lambda *shape, dt=torch.uint8: torch.empty(shape, dtype=dt, device=dev)
    kL, kR, deL, deR = mk(P, cap, 28), mk(P, cap, 28), mk(P, cap, 32), mk(P, cap, 32)
    nL, nR = torch.zeros(P, dtype=torch.int32, device=dev), torch.zeros(P, dtype=torch.int32, device=dev)
    uR, dp = mk(P, cap, dt=torch.float32), mk(P, cap, dt=torch.float32)
    ts = torch.cuda.Stream(device=dev); torch.cuda.set_stream(ts); st = ts.cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step(timed=False):
        if timed: ev[0].record()
        exL.extract_device(dL.data_ptr(), P, W, H, W, W * H, kL.data_ptr(), cap, nL.data_ptr(), deL.data_ptr(), st)
        exR.extract_device(dR.data_ptr(), P, W, H, W, W * H, kR.data_ptr(), cap, nR.data_ptr(), deR.data_ptr(), st)
        if timed: ev[1].record()
        stereo_match_device(exL, exR, P, kL.data_ptr(), deL.data_ptr(), nL.data_ptr(), kR.data_ptr(), deR.data_ptr(), nR.data_ptr(),
                            cap, c["bf"], c["fx"], uR.data_ptr(), dp.data_ptr(), st)
        if timed: ev[2].record()

    def barrier():
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup): step()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps): step()
    e1.record(); torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1); clocks = sampler.stop()
    step(timed=True); torch.cuda.synchronize()
    ms_extract, ms_match = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * P * args.steps / (ms_total * 1e-3)
    # e2e: the public batched host call orbx_stereo_extract_batch with pinned host frames in and pinned keypoints /
    # descriptors / mvuRight / mvDepth out (H2D + D2H inside the timed region, chunks pipelined over several streams)
    from orb_slam2_commit_b200 import stereo_extract_host
    pin = lambda *shape, dt=torch.uint8: torch.zeros(shape, dtype=dt).pin_memory()
    h_k = [pin(P, cap, 28) for _ in range(2)]; h_d = [pin(P, cap, 32) for _ in range(2)]
    h_n = [pin(P, dt=torch.int32) for _ in range(2)]; h_f = [pin(P, cap, dt=torch.float32) for _ in range(2)]
    hout = dict(kl=h_k[0].numpy().view(api.KP_DTYPE).reshape(P, cap), kr=h_k[1].numpy().view(api.KP_DTYPE).reshape(P, cap),
                dl=h_d[0].numpy(), dr=h_d[1].numpy(), nl=h_n[0].numpy(), nr=h_n[1].numpy(), u_right=h_f[0].numpy(), depth=h_f[1].numpy())
    outs = [h_k[0], h_d[0], h_n[0], h_k[1], h_d[1], h_n[1], h_f[0], h_f[1]]
    torch.cuda.synchronize()
    def e2e_step():
        stereo_extract_host(exL, exR, hL.numpy(), hR.numpy(), c["bf"], c["fx"], hout)
    for _ in range(2): e2e_step()
    barrier()
    n_e2e = max(3, min(args.steps, 10)); t0 = time.perf_counter()
    for _ in range(n_e2e): e2e_step()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_sync_v = world * P * n_e2e / float(t.item())
    # the same step as a stream of batches (two batches of pairs in flight, two sets of pinned result buffers)
    from orb_slam2_commit_b200 import stereo_extract_host_begin, stereo_extract_host_end
    h_k2 = [pin(P, cap, 28) for _ in range(2)]; h_d2 = [pin(P, cap, 32) for _ in range(2)]
    h_n2 = [pin(P, dt=torch.int32) for _ in range(2)]; h_f2 = [pin(P, cap, dt=torch.float32) for _ in range(2)]
    hout2 = dict(kl=h_k2[0].numpy().view(api.KP_DTYPE).reshape(P, cap), kr=h_k2[1].numpy().view(api.KP_DTYPE).reshape(P, cap),
                 dl=h_d2[0].numpy(), dr=h_d2[1].numpy(), nl=h_n2[0].numpy(), nr=h_n2[1].numpy(), u_right=h_f2[0].numpy(), depth=h_f2[1].numpy())
    houts = [hout, hout2]

    def stream_of_batches(steps):
        stereo_extract_host_begin(exL, exR, hL.numpy(), hR.numpy(), c["bf"], c["fx"], houts[0])
        for s_ in range(1, steps):
            stereo_extract_host_begin(exL, exR, hL.numpy(), hR.numpy(), c["bf"], c["fx"], houts[s_ % 2])
            stereo_extract_host_end(exL)
        stereo_extract_host_end(exL)
    stream_of_batches(2)
    barrier()
    t0 = time.perf_counter()
    stream_of_batches(n_e2e)
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_v = world * P * n_e2e / float(t.item())
    assert np.array_equal(hout["nl"], hout2["nl"]) and np.array_equal(hout["u_right"].view(np.uint32)[0, :hout["nl"][0]], hout2["u_right"].view(np.uint32)[0, :hout["nl"][0]])
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
    if rank != 0: return None
    matched = float((outs[6][:, :] >= 0).sum().item()) / P   # includes unwritten tail slots only if >= 0 garbage; informational
    nl_np = outs[2].numpy()
    matched = float(np.mean([(outs[6][i, :nl_np[i]] >= 0).sum().item() for i in range(P)]))
    line = {"metric": "stereo_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": f"{w['cfg']} stereo: {W}x{H} pairs, nFeatures={c['nfeatures']} per image, left+right extraction + ComputeStereoMatches",
                       "pairs_per_step_per_gpu": P, "distinct_pairs": len(pairs)},
            "clocks": clocks, "gpu_launches": (2 * (c["nlevels"] + 4) + 2) * args.steps,
            "e2e": {"value": e2e_v, "unit": "pairs/s", "h2d_bytes_per_step": 2 * P * W * H, "d2h_bytes_per_step": int(sum(o.numel() * o.element_size() for o in outs)), "steps": n_e2e,
                    "api": "orbx_stereo_extract_batch_begin / _end (pinned host buffers, two batches in flight)",
                    "synchronous_call": {"value": e2e_sync_v, "api": "orbx_stereo_extract_batch, one blocking call per step"}},
            "stages": {"extract_left_right_ms": ms_extract, "stereo_match_ms": ms_match},
            "pipeline": {"keypoints_per_image": float(nl_np.mean()), "stereo_matches_per_pair": matched}}
    return line


def run_stereo(args, torch, dist, rank, world, local, dev):
    w = WORKLOADS[args.workload]; c = synth.CONFIGS[w["cfg"]]
    cfgargs = (c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    line = stereo_measure(args.workload, args.batch or w["batch"], args.steps, args.warmup, torch, dist, rank, world, local, dev)
    if rank != 0: return
    pairs = [synth.synth_stereo_pair(c["width"], c["height"], 2 + i) for i in range(w["distinct"])]
    if world == 1 and not args.no_cpu_baseline:
        # CPU: the oracle port, two extractions + ComputeStereoMatches per pair, pair-parallel over the host threads
        from oracle import binding as ob
        nthreads = os.cpu_count() or 1
        exs = [(ob.Extractor(*cfgargs), ob.Extractor(*cfgargs)) for _ in range(nthreads)]
        def work(tid, iters):
            for i in range(tid, iters, nthreads):
                l, r = pairs[i % len(pairs)]
                a, b = exs[tid]
                k1, d1 = a.extract(l); k2, d2 = b.extract(r)
                ob.stereo_match(a, b, k1, d1, k2, d2, c["bf"], c["fx"])
        def run(iters):
            t0 = time.perf_counter()
            th = [threading.Thread(target=work, args=(t_, iters)) for t_ in range(nthreads)]
            [x.start() for x in th]; [x.join() for x in th]
            return time.perf_counter() - t0
        t1 = run(nthreads); iters = int(max(nthreads, nthreads * max(1.0, args.cpu_seconds / max(t1, 1e-3)))); tt = run(iters)
        line["cpu_baseline"] = {"value": iters / tt, "unit": "pairs/s", "cores": nthreads, "kind": "port", "sample": f"{iters} pairs over {nthreads} threads, {tt:.1f} s"}
    emit_json_line(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tum1", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hamming", action="store_true")
    ap.add_argument("--no-stereo", action="store_true", help="skip the EuRoC / KITTI stereo legs of the default line")
    args = ap.parse_args()
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from orb_slam2_commit_b200 import ORBextractor, api
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not os.environ.get("ORBX_NO_NUMA_BIND"):
        bind_to_gpu_numa_node(torch, local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # fd 1 already points at stderr (above), so NCCL's own banner / NCCL_DEBUG=INFO lines cannot reach the one JSON
        # line; NCCL_DEBUG is left exactly as the launcher set it
        dist.init_process_group("nccl", device_id=dev)

    if any(WORKLOADS[args.workload].get(k) for k in ("stereo", "frame", "track", "matchers")):
        wl = WORKLOADS[args.workload]
        (run_stereo if wl.get("stereo") else run_frame if wl.get("frame") else run_track if wl.get("track") else run_matchers)(
            args, torch, dist, rank, world, local, dev)
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return
    w = WORKLOADS[args.workload]
    c = synth.CONFIGS[w["cfg"]]
    B = args.batch or w["batch"]
    W, H = c["width"], c["height"]
    # every rank extracts its own shard of the frame stream (frame index -> rank); seeds differ per rank
    frames = make_frames(c, w["distinct"], seed0=1 + 1000 * rank)
    host_batch = torch.empty((B, H, W), dtype=torch.uint8, pin_memory=True)
    hb = host_batch.numpy()
    for i in range(B):
        hb[i] = frames[i % len(frames)]
    d_imgs = host_batch.to(dev)

    ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"], device=local)
    rectify = bool(w.get("rectify"))
    if rectify:
        rmaps = synth.rectify_maps(W, H)
        ex.set_rectify_maps(*rmaps)
    cap = ex.reserve(W, H, B)
    d_kps = torch.empty((B, cap, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
    d_nkp = torch.zeros(B, dtype=torch.int32, device=dev)
    # a dedicated (non-default) stream: the kernels, the timing events and the torch events all live on it
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    def step():
        if rectify:
            ex.extract_device_rectified(d_imgs.data_ptr(), B, W, W * H, d_kps.data_ptr(), cap, d_nkp.data_ptr(), d_desc.data_ptr(), stream)
        else:
            ex.extract_device(d_imgs.data_ptr(), B, W, H, W, W * H, d_kps.data_ptr(), cap, d_nkp.data_ptr(), d_desc.data_ptr(), stream)
    host_call = ex.extract_host_rectified if rectify else ex.extract_host

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    ex.enable_timing(True)
    sampler = ClockSampler(local); sampler.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    stage_ms, nruns = ex.stage_ms()
    ex.enable_timing(False)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    nkp = d_nkp.cpu().numpy()
    nkp_mean = float(nkp.mean())
    frames_per_s = world * B * args.steps / (ms_total * 1e-3)

    # ---- e2e through the public C-ABI call with pinned host buffers
    h_kps = torch.empty((B, cap, 28), dtype=torch.uint8, pin_memory=True)
    h_desc = torch.empty((B, cap, 32), dtype=torch.uint8, pin_memory=True)
    h_nkp = torch.zeros(B, dtype=torch.int32, pin_memory=True)
    kps_np = h_kps.numpy().view(api.KP_DTYPE).reshape(B, cap)
    desc_np = h_desc.numpy(); nkp_np = h_nkp.numpy()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        host_call(hb, kps_np, desc_np, nkp_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_call(hb, kps_np, desc_np, nkp_np)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_sync_fps = world * B * e2e_steps / e2e_s
    assert int(nkp_np.sum()) == int(nkp.sum()), "host path and device path disagree"
    # the same step as a stream of batches: orbx_extract_batch_begin / _end with two batches in flight (two sets of pinned
    # result buffers), so the upload of step i+1 overlaps the kernels of step i; every step's H2D and D2H stay in the timed region
    e2e_fps, e2e_api = e2e_sync_fps, ("orbx_extract_batch_rectified" if rectify else "orbx_extract_batch") + " (pinned host buffers)"
    if not rectify:
        h_kps2 = torch.empty((B, cap, 28), dtype=torch.uint8, pin_memory=True)
        h_desc2 = torch.empty((B, cap, 32), dtype=torch.uint8, pin_memory=True)
        h_nkp2 = torch.zeros(B, dtype=torch.int32, pin_memory=True)
        sets = [(kps_np, desc_np, nkp_np), (h_kps2.numpy().view(api.KP_DTYPE).reshape(B, cap), h_desc2.numpy(), h_nkp2.numpy())]

        def stream_of_batches(steps):
            ex.extract_host_begin(hb, *sets[0])
            for s_ in range(1, steps):
                ex.extract_host_begin(hb, *sets[s_ % 2])
                ex.extract_host_end()
            ex.extract_host_end()
        stream_of_batches(2)
        barrier()
        t0 = time.perf_counter()
        stream_of_batches(e2e_steps)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_fps = world * B * e2e_steps / float(t.item())
        e2e_api = "orbx_extract_batch_begin / _end (pinned host buffers, two batches in flight)"
        assert int(sets[0][2].sum()) == int(nkp.sum()) and int(sets[1][2].sum()) == int(nkp.sum()), "batches in flight disagree with the device path"

    # ---- single-frame latency through orbx_extract (what a SLAM front-end sees: one frame in, keypoints out)
    lat_ms = lat_pyr_ms = None
    if rank == 0 and not rectify:
        ex1 = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"], device=local)
        cap1 = ex1.reserve(W, H, 1)
        k1 = h_kps.numpy().view(api.KP_DTYPE).reshape(-1)[:cap1].reshape(1, cap1)
        d1 = desc_np.reshape(-1, 32)[:cap1].reshape(1, cap1, 32)
        n1 = nkp_np[:1]
        for _ in range(20):
            ex1.extract_host(hb[:1], k1, d1, n1)
        t0 = time.perf_counter()
        for i in range(200):
            ex1.extract_host(hb[i % B:i % B + 1], k1, d1, n1)
        lat_ms = (time.perf_counter() - t0) / 200 * 1e3
        # the same call when the caller also wants mvImagePyramid on the host (stereo: Frame.cc:556,681-700): the frame's raw
        # pyramid block comes back in ONE asynchronous copy in the stream of the kernels (orbx_set_pyramid_mirror)
        ex1.set_pyramid_mirror(True)
        for _ in range(20):
            ex1.extract_host(hb[:1], k1, d1, n1)
        t0 = time.perf_counter()
        for i in range(200):
            ex1.extract_host(hb[i % B:i % B + 1], k1, d1, n1)
            ex1.pyramid_mirror(0)
        lat_pyr_ms = (time.perf_counter() - t0) / 200 * 1e3
        del ex1

    # ---- what the box can feed: all ranks upload their pinned batch at once, no kernels (the e2e number is read against this)
    ceiling_gbs, duplex_gbs = h2d_ceiling_gbs(torch, dist, host_batch, dev, world, d2h_bytes=B * (cap * 60 + 4))

    # ---- configs 2 and 3 ride along on the default line: KITTI and EuRoC stereo pairs (two extractors + ComputeStereoMatches,
    # Frame.cc:80-84, 547-788), pairs sharded by frame over the ranks, short runs
    stereo_legs = {}
    if args.workload == "tum1" and not args.no_stereo:
        torch.cuda.set_stream(torch.cuda.default_stream(dev))
        for leg, P_leg in (("euroc_stereo", 128), ("kitti_stereo", 64)):
            ln = stereo_measure(leg, P_leg, max(3, min(args.steps, 8)), 3, torch, dist, rank, world, local, dev, distinct=4)
            if ln is not None:
                stereo_legs[leg] = {"pairs_per_s": ln["value"], "unit": "pairs/s", "ms_per_step": ln["ms_per_step"], "e2e": ln["e2e"],
                                    "config": ln["config"], "stages": ln["stages"], "pipeline": ln["pipeline"]}
        torch.cuda.set_stream(tstream)

    # ---- config 4 across ranks: train set sharded, per-query top-2 merged after an NCCL all-gather (all ranks take part)
    hamming_sharded = None
    if world > 1 and not args.no_hamming:
        from orb_slam2_commit_b200 import dist as od
        nq, nt = 2048, 1_000_000
        a, b = od.shard_range(nt, world, rank)
        # the seeded config-4 set (duplicated train rows -> exact ties, queries = train rows with 0..40 flipped bits):
        # every rank generates the same arrays and keeps its contiguous shard of the train rows
        train_np, query_np = synth.synth_descriptors(nt, nq, seed=42)
        shard = torch.from_numpy(train_np[a:b]).to(dev)
        query = torch.from_numpy(query_np).to(dev)
        for _ in range(3):
            od.hamming_top2_sharded(query, shard, a)
        barrier()
        h0 = torch.cuda.Event(enable_timing=True); h1 = torch.cuda.Event(enable_timing=True)
        reps = 10
        h0.record()
        for _ in range(reps):
            od.hamming_top2_sharded(query, shard, a)
        h1.record(); torch.cuda.synchronize()
        t = torch.tensor([h0.elapsed_time(h1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # same workload with the exchange fused into the matcher (peer stores over NVLink + arrival counters, no NCCL call)
        pm = od.PeerHammingMatcher(nq)
        for _ in range(3):
            pm(query, shard, a)
        barrier()
        h0.record()
        for _ in range(reps):
            pm(query, shard, a, check=False)          # the status word is read once after the loop
        h1.record(); torch.cuda.synchronize()
        pm.raise_if_failed()
        t = torch.tensor([h0.elapsed_time(h1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_fused = float(t.item())
        res_peer = pm(query, shard, a); res_nccl = od.hamming_top2_sharded(query, shard, a)
        same = all(bool(torch.equal(x, y)) for x, y in zip(res_peer, res_nccl))
        peer_ok = int(pm.status.item()) == 0 and same
        # parity outside the timed region: the whole train set on ONE GPU through orbx_hamming_top2_device (rank 0), against
        # what each exchange path returned on every rank (ORBmatcher.cc:84-126 semantics: first index wins, strict <)
        par = torch.zeros(2, dtype=torch.int32, device=dev)
        if rank == 0:
            full = torch.from_numpy(train_np).to(dev)
            single = od.hamming_top2_single(query, full)
            ref_t = torch.stack(single).contiguous()
        else:
            ref_t = torch.empty((3, nq), dtype=torch.int32, device=dev)
        dist.broadcast(ref_t, 0)
        par[0] = int(all(bool(torch.equal(x, ref_t[i])) for i, x in enumerate(res_nccl)))
        par[1] = int(all(bool(torch.equal(x, ref_t[i])) for i, x in enumerate(res_peer)))
        dist.all_reduce(par, op=dist.ReduceOp.MIN)
        parity_nccl, parity_peer = bool(par[0].item()), bool(par[1].item())
        barrier()
        pm.close()
        hamming_sharded = {"workload": f"2048 queries x 1,000,000 train rows sharded over {world} GPUs",
                           "nccl_allgather_merge": {"ms": ms, "matches_per_s": nq / (ms * 1e-3), "pair_distances_per_s": nq * nt / (ms * 1e-3)},
                           "fused_peer_stores": {"ms": ms_fused, "matches_per_s": nq / (ms_fused * 1e-3),
                                                 "pair_distances_per_s": nq * nt / (ms_fused * 1e-3), "equal_to_nccl_path": peer_ok},
                           "parity_vs_single_gpu": parity_nccl and parity_peer,
                           "parity_detail": {"nccl_allgather_merge": parity_nccl, "fused_peer_stores": parity_peer,
                                             "checked": "idx1 / dist1 / dist2 of all 2048 queries on every rank against orbx_hamming_top2_device over the whole 1,000,000-row train set on rank 0",
                                             "data": "synth.synth_descriptors(1000000, 2048, seed=42): duplicated train rows (exact ties), queries 0..40 bit flips away from a train row"}}

    if rank == 0:
        peaks = {"hbm_gbs": 6650.0, "source": "fallback"}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            try:
                peaks = {"hbm_gbs": float(json.load(open(pk))["hbm_gbs"]), "source": "measured"}
            except Exception:
                pass
        ab = algorithmic_bytes(c, nkp_mean)
        names = ["pyramid", "fast", "quadtree", "describe"]
        stages = {n: {"ms_per_launch_group": float(stage_ms[i]), "share": float(stage_ms[i] / max(stage_ms.sum(), 1e-9))} for i, n in enumerate(names)}
        dom = int(np.argmax(stage_ms))
        dom_bytes = ab[names[dom]] * B
        achieved = dom_bytes / (float(stage_ms[dom]) * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture (profiles/), scaled from the
        # captured batch (128 frames per launch) to this run's frames per launch; None when no capture is present
        traffic = None
        try:
            import csv as _csv
            import glob as _glob
            # the newest committed capture of THIS geometry: (file suffix, frames per captured launch)
            cap_name, cap_frames = {"tum1": ("batch128", 128), "kitti": ("kitti", 128), "euroc": ("euroc", 128), "4k": ("4k", 16)}.get(w["cfg"], (None, 0))
            prof = sorted(_glob.glob(os.path.join(ROOT, "profiles", f"r0*_ncu_full_{cap_name}.csv")))[-1]
            prof_name = os.path.relpath(prof, ROOT)
            key = {"pyramid": "pyr_resize_kernel", "fast": "fast_cells_kernel", "quadtree": "quadtree_kernel", "describe": "describe_kernel"}[names[dom]]
            rows_ = list(_csv.reader(open(prof)))
            hdr_ = rows_[0]
            ir = [i for i, c_ in enumerate(hdr_) if c_.startswith("dram__bytes_read.sum")][0]
            iw = [i for i, c_ in enumerate(hdr_) if c_.startswith("dram__bytes_write.sum")][0]
            unit_ = lambda c_: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[c_[c_.index("[") + 1:c_.index("]")]]
            by = sum(float(r_[ir]) * unit_(hdr_[ir]) + float(r_[iw]) * unit_(hdr_[iw]) for r_ in rows_[1:] if key in r_[0])
            # a capture of one geometry says nothing about another workload's traffic (nor about the fused-remap level 0)
            traffic = by / cap_frames * B if not rectify else None
        except Exception:
            traffic = None; prof_name = None; cap_frames = 0
        # the binding resource of this integer/byte pipeline is instruction issue, not HBM: warp-instructions per frame
        # (committed ncu capture, TUM1 geometry) x measured frames/s against 148 SMs x 4 issue slots x the sampled SM clock
        issue = None
        try:
            ii = [i for i, c_ in enumerate(hdr_) if c_.startswith("smsp__inst_executed.sum")][0]
            inst_per_frame = sum(float(r_[ii]) for r_ in rows_[1:]) / cap_frames
            if cap_frames and not rectify:
                peak_issue = 148 * 4 * float(clocks.get("sm_mhz") or 1965.0) * 1e6
                issue = {"warp_instructions_per_frame": inst_per_frame, "peak_warp_instructions_per_s": peak_issue,
                         "frac": inst_per_frame * frames_per_s / world / peak_issue,
                         "source": f"{prof_name} (smsp__inst_executed.sum, all kernels of a step)"}
        except Exception:
            issue = None
        # what binds: the pixel kernels of this integer / byte path saturate instruction issue (ALU pipe) or, for describe,
        # shared-memory wavefronts long before HBM (ncu: issue 73-80 %, DRAM 4-7 %); `achieved` / `peak` / `frac` keep the HBM
        # roofline of the contract beside it
        roofline = {"bound": "issue", "bound_detail": "instruction issue / ALU pipe (ncu smsp__issue_active ~78 %, dram throughput ~5 %); achieved, peak and frac are the HBM roofline of the same kernel",
                    "issue": issue, "kernel": {"pyramid": "pyr_level0_kernel+pyr_resize_kernel (one launch per level)", "fast": "fast_cells_kernel",
                                              "quadtree": "quadtree_kernel", "describe": "blur_units_kernel+describe_kernel"}[names[dom]],
                    "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "peak_source": peaks["source"],
                    "algorithmic_bytes_per_frame": ab[names[dom]], "frames_per_launch": B, "traffic": traffic,
                    "traffic_source": (f"{prof_name} (dram read+write per frame x frames per launch)" if traffic is not None else None),
                    "stage_events_averaged_over_steps": nruns}
        launches_per_step = c["nlevels"] + 4          # one per pyramid level, FAST, quadtree, blur, describe
        line = {"metric": "orb_frames_per_s", "value": frames_per_s, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": f"{w['cfg']}{' + fused cv::remap rectification' if rectify else ''}: {W}x{H} mono frames, nFeatures={c['nfeatures']}, scale {c['scale']}, {c['nlevels']} levels, FAST {c['ini_th']}/{c['min_th']}",
                           "frames_per_step_per_gpu": B, "distinct_frames": len(frames), "sharding": "by frame, no collective",
                           "l2": f"batch working set {B * (W * H + ab['A'] * 1.3) / 1e6:.0f} MB per step > 126 MB L2 (inputs {B * W * H / 1e6:.0f} MB)"},
                "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
                "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": B * W * H,
                        "d2h_bytes_per_step": B * (cap * 60 + 4), "steps": e2e_steps, "api": e2e_api,
                        "synchronous_call": {"value": e2e_sync_fps, "api": ("orbx_extract_batch_rectified" if rectify else "orbx_extract_batch") + ", one blocking call per step"},
                        "h2d_ceiling_gbs": ceiling_gbs, "h2d_gbs": e2e_fps * W * H / 1e9, "frac_of_h2d_ceiling": e2e_fps * W * H / 1e9 / ceiling_gbs,
                        "h2d_ceiling_how": "all ranks copy their pinned input batch to their GPU concurrently, no kernels running (aggregate GB/s, max time over ranks)",
                        "duplex_ceiling_gbs": duplex_gbs, "duplex_gbs": e2e_fps * (W * H + cap * 60 + 4) / 1e9,
                        "frac_of_duplex_ceiling": e2e_fps * (W * H + cap * 60 + 4) / 1e9 / duplex_gbs,
                        "duplex_ceiling_how": "the same with a second stream copying the step's result bytes (keypoints + descriptors + counts) back to pinned memory at the same time; both directions counted"},
                "latency_single_frame_ms": lat_ms, "latency_single_frame_with_pyramid_ms": lat_pyr_ms,
                "roofline": roofline, "stages": stages,
                "pipeline": {"keypoints_per_frame": nkp_mean, "keypoints_per_s": frames_per_s * nkp_mean,
                             "algorithmic_bytes_per_frame_survey": ab["B_survey"],
                             "hbm_frac_of_survey_bytes": frames_per_s / world * ab["B_survey"] / 1e9 / peaks["hbm_gbs"]}}
        if not args.no_hamming and world == 1:
            try:
                line["hamming"] = bench_hamming(torch, api.lib(), dev, peaks, clocks.get("sm_mhz"))
            except Exception as e:  # never lose the main line
                line["hamming"] = {"error": str(e)}
        if hamming_sharded is not None:
            line["hamming"] = hamming_sharded
        line.update(stereo_legs)
        if world == 1 and not args.no_cpu_baseline:
            nthreads = os.cpu_count() or 1
            if rectify:
                from oracle import binding as ob
                def make_worker(tid):
                    e = ob.RefExtractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"]) if ob.ref_available() else \
                        ob.Extractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
                    return lambda i: e.extract(ob.remap(frames[i % 8], *rmaps))
                fps, sample = cpu_thread_bench(make_worker, args.cpu_seconds, nthreads)
                kind, kpf = ("reference" if ob.ref_available() else "port"), None
                sample += " (oracle remap + extractor per frame)"
            else:
                fps, kind, sample, kpf = cpu_reference_bench(c, frames[:8], args.cpu_seconds, nthreads)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": nthreads, "kind": kind, "sample": sample,
                                    "keypoints_per_frame": kpf}
        else:
            line["cpu_baseline"] = None
        emit_json_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
