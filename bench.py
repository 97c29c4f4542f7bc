#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native ORB front-end (BASELINE.json configs[1]).

Workload (one "step"): a batch of synthetic 640x480 frames -> ORB extraction (1000 features, 8 levels, 1.2, FAST
20/7) of every frame + frame-to-frame Hamming kNN (k=2, TH_LOW 50, ratio 0.7) between consecutive frames.
Frames are independent units.  Default (`--scaling strong`, the config as BASELINE.json writes it): ONE 4096-frame batch is
block-partitioned over the N GPUs, every rank extracts its block + the replicated boundary frame and matches its own pairs
(no data-path collective); `value` is whole-job frames/s.  `--scaling weak` gives every GPU its own 4096-frame batch; at N > 1
the other mode is measured beside the headline one (`other_scaling`).

  python bench.py --gpus N --steps K --warmup W            # B200 arm (one process per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: the reference's own CPU extractor
                                                           # (oracle/_ref) + CPU matcher on all host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions of value / e2e / roofline /
cpu_baseline.
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

W_IMG, H_IMG = 640, 480
NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH = 1000, 1.2, 8, 20, 7
TH_LOW, RATIO = 50, 0.7
METRIC = "ORB frames/s @640x480/1000kp (extract + frame-to-frame Hamming kNN k=2, ratio 0.7)"


def level_sizes():
    sc = [1.0]
    for _ in range(1, NLEVELS):
        sc.append(float(np.float32(sc[-1] * float(np.float32(SCALE)))))
    out = []
    for s in sc:
        inv = np.float32(1.0) / np.float32(s)
        out.append((int(np.rint(np.float32(W_IMG) * inv)), int(np.rint(np.float32(H_IMG) * inv))))
    return out


def algorithmic_bytes():
    """SURVEY.md §8(d), stage-materialised convention, per frame."""
    sz = level_sizes()
    A = sum(w * h for w, h in sz); A0 = sz[0][0] * sz[0][1]; A7 = sz[-1][0] * sz[-1][1]
    return {"pyramid": (A - A7) + (A - A0), "fast": A, "quadtree": 4 * 21000, "blur": 2 * A, "orient_desc": 60 * NFEAT,
            "frame_total": 5 * A - A0 - A7 + 60 * NFEAT}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index`, so that the pinned staging buffers of the
    end-to-end path are allocated (first touch) on the NUMA node the GPU hangs off.  Best effort; returns what was verified:
    {"cpus": n bound, "node": NUMA node of the first bound CPU (from /sys), "nodes_online": ...} or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1 and 64 * i + b < ncpu]
        if cpus:
            os.sched_setaffinity(0, cpus)
        node = None
        for n in sorted(os.listdir("/sys/devices/system/node")):
            if n.startswith("node") and os.path.exists("/sys/devices/system/node/%s/cpu%d" % (n, cpus[0] if cpus else 0)):
                node = int(n[4:])
        online = open("/sys/devices/system/node/online").read().strip()
        return {"cpus": len(cpus), "node": node, "nodes_online": online}
    except Exception:
        return None


_CPU_MATCHER = {}


def cpu_matcher(port, march=None):
    """The matcher leg of the CPU arm: the reference's own computeDistance (matcher.cpp compiled in place, oracle/_ref) inside its
    matching loop; the CPU port of the same loop where that library is absent."""
    if march not in _CPU_MATCHER:
        import oracle
        try:
            _CPU_MATCHER[march] = oracle.RefMatcher(march=march)
        except Exception:
            _CPU_MATCHER[march] = port
    return _CPU_MATCHER[march]


def cpu_reference_step(imgs, cores, Ref, port, march=None):
    """The reference's own extractor (oracle/_ref, one instance per core over disjoint frames, ORBextractor.h:85 is
    stateful) + the reference's matcher loop (threaded over queries) on `imgs`.  Returns seconds."""
    m = cpu_matcher(port, march)
    t0 = time.perf_counter()
    kps, desc, cnt = Ref.extract_batch(imgs, cores, keep_outputs=True)
    for f in range(len(imgs) - 1):
        m.knn2(desc[f, :cnt[f]], desc[f + 1, :cnt[f + 1]], TH_LOW, RATIO, nthreads=cores)
    return time.perf_counter() - t0


def cv2_primitives_ms(img):
    """Single-thread time of the REAL OpenCV primitives (the cv2 wheel, SIMD code) for one frame's pyramid resizes, per-cell
    FAST calls and blurs, next to the scalar restatements the compiled reference links.  None if cv2 is
    not importable on this box."""
    try:
        import cv2
    except Exception:       # noqa: BLE001
        return None
    cv2.setNumThreads(1)
    sizes = level_sizes()

    def once():
        t = {}
        t0 = time.perf_counter()
        pyr = [img]
        for (w, h) in sizes[1:]:
            pyr.append(cv2.resize(pyr[-1], (w, h), interpolation=cv2.INTER_LINEAR))
        t["resize"] = time.perf_counter() - t0
        det = {th: cv2.FastFeatureDetector_create(th, True) for th in (INI_TH, MIN_TH)}
        t0 = time.perf_counter()
        calls = 0
        for lv in pyr:
            H, W = lv.shape
            x0, y0, x1, y1 = 16 - 3, 16 - 3, W - 16 + 3, H - 16 + 3
            wd, hg = x1 - x0, y1 - y0
            nc, nr = max(wd // 30, 1), max(hg // 30, 1)
            wc, hc = -(-wd // nc), -(-hg // nr)
            for i in range(nr):
                iy = y0 + i * hc
                if iy >= y1 - 3:
                    continue
                my = min(iy + hc + 6, y1)
                for j in range(nc):
                    ix = x0 + j * wc
                    if ix >= x1 - 6:
                        continue
                    mx = min(ix + wc + 6, x1)
                    roi = lv[iy:my, ix:mx]
                    calls += 1
                    if not det[INI_TH].detect(roi):
                        det[MIN_TH].detect(roi)
        t["fast_cells"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        for lv in pyr:
            cv2.GaussianBlur(lv, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        t["blur"] = time.perf_counter() - t0
        return t, calls
    once()
    best, calls = None, 0
    for _ in range(3):
        t, calls = once()
        if best is None or sum(t.values()) < sum(best.values()):
            best = t
    out = {k: v * 1e3 for k, v in best.items()}
    out["total"] = sum(out.values())
    out["fast_python_calls"] = calls
    out["threads"] = 1
    return out


def cpu_arm(imgs, cores, steps=1, warm=True):
    """The CPU arm on `imgs` (a bounded sample of the bench sequence): (c) the reference's own sources with the reference's flags
    (-O3, baseline x86-64: CMakeLists.txt:4-5) -- the number `vs_reference` uses -- plus (a) the same sources at the highest
    x86-64 level this host supports (the '-march=native' stand-in, oracle/Makefile) and (b) single-thread timings of the real cv2
    primitives.  Returns (value frames/s, kind, extras)."""
    import oracle
    port = oracle.Port(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH)
    kind = "reference"
    try:
        Ref = oracle.Ref(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, parity=False)
    except Exception:
        kind = "port"       # no prebuilt oracle/_ref on this box: the CPU port of the same algorithm

        class _P:
            def extract_batch(self, im, n, keep_outputs=True):
                return port.extract_batch(im, n)
        Ref = _P()
    if warm:
        cpu_reference_step(imgs[:max(cores, 2)], cores, Ref, port)
    t = 0.0
    for _ in range(steps):
        t += cpu_reference_step(imgs, cores, Ref, port)
    value = len(imgs) * steps / t
    extras = {"reference_flags": {"value": value, "unit": "frames/s", "flags": "-O3 (baseline x86-64, CMakeLists.txt:4-5)", "cores": cores}}
    try:
        march = oracle.best_march() if kind == "reference" else None
        if march:
            RefN = oracle.Ref(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, parity=False, march=march)
            cpu_reference_step(imgs[:max(cores, 2)], cores, RefN, port, march)
            tn = cpu_reference_step(imgs, cores, RefN, port, march)
            extras["march_native"] = {"value": len(imgs) / tn, "unit": "frames/s", "cores": cores,
                                      "flags": "-O3 -march=x86-64-%s (highest level this host's /proc/cpuinfo supports)" % march}
    except Exception as e:      # noqa: BLE001
        extras["march_native"] = {"error": str(e)[:100]}
    try:
        extras["cv2_primitives_ms"] = cv2_primitives_ms(np.ascontiguousarray(imgs[0]))
    except Exception as e:      # noqa: BLE001
        extras["cv2_primitives_ms"] = {"error": str(e)[:100]}
    return value, kind, t, extras


def run_reference(args, rank, world):
    """--impl reference: CPU arm.  Rank 0 alone works; the other ranks exit 0."""
    if rank != 0:
        return
    from vo_slam_test_b200 import synth
    cores = os.cpu_count() or 1
    sample = max(cores * 8, 16)
    imgs = synth.make_sequence(sample, seed=0)
    value, kind, t, extras = cpu_arm(imgs, cores, steps=args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * t / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "extract+match, bounded sample of %d synthetic 640x480 frames per step (of the 4096-frame "
                                   "batch), 1000 features, 8 levels, 1.2, FAST 20/7, kNN k=2 TH_LOW 50 ratio 0.7" % sample,
                       "frames_per_step": sample},
            "cpu_baseline": dict({"value": value, "unit": "frames/s", "cores": cores, "kind": kind,
                                  "sample": "%d frames x %d steps, one extractor per core + threaded matcher" % (sample, args.steps)},
                                 **extras),
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


class Block:
    """One rank's share of a step: frames [lo, hi_ext) of the sequence (its block + the replicated boundary frame) and the pairs
    (p, p+1), p in [lo, p_hi) -- device-resident copies, output buffers, pinned host mirrors for the end-to-end path."""

    def __init__(self, torch, ex, dev, synth, seed, lo, hi, hi_ext, p_hi):
        self.torch, self.ex, self.dev = torch, ex, dev
        self.lo, self.hi, self.hi_ext, self.p_hi = lo, hi, hi_ext, p_hi
        self.F = F = hi_ext - lo
        self.owned = hi - lo
        self.npairs = npairs = p_hi - lo
        self.cap = cap = ex.max_keypoints
        self.wc = None
        if os.environ.get("ORBX_BENCH_WC", "0") == "1":       # experiment: write-combined pinned input (orbx_host_alloc)
            from vo_slam_test_b200 import api
            self.wc = api.HostBuffer((F, H_IMG, W_IMG), write_combined=True)
            tmp = synth.make_sequence(F, seed=seed, start=lo)
            self.wc.array[...] = tmp
            self.h_imgs = torch.from_numpy(self.wc.array)
        else:
            self.h_imgs = torch.empty((F, H_IMG, W_IMG), dtype=torch.uint8, pin_memory=True)
            synth.make_sequence(F, seed=seed, out=self.h_imgs.numpy(), start=lo)
        self.d_imgs = self.h_imgs.to(dev)
        self.d_kps = torch.empty((F, cap, 7), dtype=torch.float32, device=dev)
        self.d_desc = torch.empty((F, cap, 32), dtype=torch.uint8, device=dev)
        self.d_counts = torch.zeros(F, dtype=torch.int32, device=dev)
        self.d_qf = torch.arange(0, max(npairs, 1), dtype=torch.int32, device=dev)
        self.d_tf = self.d_qf + 1
        np1 = max(npairs, 1)
        self.d_midx = torch.empty((np1, cap), dtype=torch.int32, device=dev)
        self.d_md1 = torch.empty_like(self.d_midx); self.d_md2 = torch.empty_like(self.d_midx)
        self.d_mok = torch.zeros((np1, cap), dtype=torch.uint8, device=dev)
        self.host = None

    def step_resident(self, stream):
        """One pass of the hot path over the resident batch: extraction of the whole batch on two lanes, then ONE matching launch.
        ORBX_BENCH_PIPELINED_MATCH=1: orbx_extract_match_batch_device instead (the top-2 of every chunk launched on its lane right
        behind the extraction; measured equal: 202.5 k against 203.1 k frames/s, every stage is bound by the same integer pipe)."""
        from vo_slam_test_b200 import api
        if self.npairs > 0 and os.environ.get("ORBX_BENCH_PIPELINED_MATCH", "0") == "1":
            self.ex.extract_match_batch_device(self.d_imgs.data_ptr(), self.F, W_IMG, H_IMG, W_IMG, W_IMG * H_IMG, self.d_kps.data_ptr(),
                                               self.d_desc.data_ptr(), self.cap, self.d_counts.data_ptr(), TH_LOW, RATIO,
                                               self.d_midx.data_ptr(), self.d_md1.data_ptr(), self.d_md2.data_ptr(),
                                               self.d_mok.data_ptr(), stream)
            return
        self.ex.extract_batch_device(self.d_imgs.data_ptr(), self.F, W_IMG, H_IMG, W_IMG, W_IMG * H_IMG, self.d_kps.data_ptr(),
                                     self.d_desc.data_ptr(), self.cap, self.d_counts.data_ptr(), stream)
        if self.npairs > 0:
            api.knn2_pairs_device(self.d_desc.data_ptr(), self.d_counts.data_ptr(), self.cap, self.d_qf.data_ptr(), self.d_tf.data_ptr(),
                                  self.npairs, TH_LOW, RATIO, self.d_midx.data_ptr(), self.d_md1.data_ptr(), self.d_md2.data_ptr(),
                                  self.d_mok.data_ptr(), stream)

    def alloc_host(self):
        torch, F, cap, np1 = self.torch, self.F, self.cap, max(self.F - 1, 1)
        self.host = {"kps": torch.empty((F, cap, 7), dtype=torch.float32, pin_memory=True),
                     "desc": torch.empty((F, cap, 32), dtype=torch.uint8, pin_memory=True),
                     "counts": torch.empty(F, dtype=torch.int32, pin_memory=True),
                     "midx": torch.empty((np1, cap), dtype=torch.int32, pin_memory=True),
                     "md1": torch.empty((np1, cap), dtype=torch.int32, pin_memory=True),
                     "md2": torch.empty((np1, cap), dtype=torch.int32, pin_memory=True),
                     "mok": torch.empty((np1, cap), dtype=torch.uint8, pin_memory=True)}

    def step_e2e(self):
        """The public host entry point: pinned host frames in, pinned host results out (all F-1 consecutive pairs of the block;
        with the halo frame last that is exactly the rank's own pairs -- the last rank has no halo and one pair less)."""
        h = self.host
        self.ex.extract_match_batch(self.h_imgs.data_ptr(), self.F, W_IMG, H_IMG, h["kps"].data_ptr(), h["desc"].data_ptr(), self.cap,
                                    h["counts"].data_ptr(), TH_LOW, RATIO, h["midx"].data_ptr(), h["md1"].data_ptr(),
                                    h["md2"].data_ptr(), h["mok"].data_ptr())

    def bytes_e2e(self):
        F, cap = self.F, self.cap
        return F * W_IMG * H_IMG, F * 4 + F * cap * 28 + F * cap * 32 + max(F - 1, 0) * cap * 13

    def checksum(self):
        """Integer digests of what this rank OWNS (halo frame excluded): equal sums at every N prove the sharded job produced the
        N = 1 results (rank 0 all-reduces them)."""
        torch = self.torch
        n = self.owned
        cnt = self.d_counts[:n].to(torch.int64).clamp(max=self.cap)
        rows = torch.arange(self.cap, device=self.dev)[None, :] < cnt[:, None]
        dsum = (self.d_desc[:n].to(torch.int64).sum(dim=2) * rows).sum()
        out = [cnt.sum(), dsum]
        if self.npairs > 0:
            ok = self.d_mok[:self.npairs].to(torch.int64)
            qrows = rows[:self.npairs]
            out += [(ok * qrows).sum(), ((self.d_midx[:self.npairs].to(torch.int64) + 1) * ok * qrows).sum(),
                    (self.d_md1[:self.npairs].to(torch.int64) * qrows).sum()]
        else:
            z = torch.zeros((), dtype=torch.int64, device=self.dev)
            out += [z, z, z]
        return torch.stack(out)


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import vo_slam_test_b200 as vo
    from vo_slam_test_b200 import api, synth, sharded

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = None
    if world > 1:
        numa = bind_to_gpu_numa_node(local_rank)     # before any pinned allocation: first touch on the GPU's NUMA node
    ex = vo.ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, device=local_rank)
    cap = ex.max_keypoints
    stream = torch.cuda.current_stream().cuda_stream
    steps, warm = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure(blk, job_frames, sample_clocks):
        """warm-up, K resident steps (CUDA events on the launching stream, max over ranks), then the same through the host API."""
        for _ in range(warm):
            blk.step_resident(stream)
        torch.cuda.synchronize()
        launches0 = ex.launch_count() + vo.lib().hamm_launch_count()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            blk.step_resident(stream)
        e1.record()
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop() if sampler else None
        launches = ex.launch_count() + vo.lib().hamm_launch_count() - launches0
        res = {"value": job_frames * steps / (ms_total / 1000.0), "ms_per_step": ms_total / steps, "launches": int(launches), "clocks": clocks}
        res["checksum"] = blk.checksum()
        counts = blk.d_counts.cpu().numpy()
        # end to end through the host C ABI
        blk.alloc_host()
        blk.step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            blk.step_e2e()
        torch.cuda.synchronize()
        te = max_over_ranks(time.perf_counter() - t0)
        assert np.array_equal(blk.host["counts"].numpy(), counts), "e2e and resident paths disagree"
        h2d, d2h = blk.bytes_e2e()
        res["e2e"] = {"value": job_frames * steps / te, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                      "ms_per_step": 1000.0 * te / steps}
        res["counts"] = counts
        return res

    # ---- the headline workload: BASELINE configs[1] ------------------------------------------------------------
    Ftot = args.frames
    if args.scaling == "strong":
        lo, hi, hi_ext, p_hi = sharded.strong_block(Ftot, rank, world)
        blk = Block(torch, ex, dev, synth, 0, lo, hi, hi_ext, p_hi)
        job_frames = Ftot
    else:
        blk = Block(torch, ex, dev, synth, rank, 0, Ftot, Ftot, Ftot - 1)
        job_frames = world * Ftot
    main = measure(blk, job_frames, True)
    value, ms_step, clocks, launches, counts = main["value"], main["ms_per_step"], main["clocks"], main["launches"], main["counts"]
    chk = main["checksum"]
    if world > 1:
        dist.all_reduce(chk)
    chk = [int(v) for v in chk.cpu().tolist()]
    owned_pairs = (Ftot - 1) if args.scaling == "strong" else world * (Ftot - 1)
    checksum = {"keypoints": chk[0], "descriptor_byte_sum": chk[1], "accepted_matches": chk[2], "accepted_index_sum": chk[3],
                "best_distance_sum": chk[4],
                "note": "sums over every frame / pair of the job, each counted by the rank that owns it (halo frames excluded); "
                        "identical at N = 1, 2, 4, 8 under strong scaling"}
    F = blk.F
    npairs = blk.npairs
    h2d, d2h = main["e2e"]["h2d_bytes_per_step"], main["e2e"]["d2h_bytes_per_step"]
    e2e_value = main["e2e"]["value"]

    # ---- parity spot check against the CPU oracle on rank 0 (every N): 4 frames + 3 pairs of the timed outputs ----------------
    parity = None
    if rank == 0 and not args.skip_cpu:
        import oracle
        port = oracle.Port(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH)
        nchk = min(4, F)
        rk, rd, rc = port.extract_batch(blk.h_imgs.numpy()[:nchk], min(os.cpu_count() or 1, 4))
        hk = blk.d_kps[:nchk].cpu().numpy(); hd = blk.d_desc[:nchk].cpu().numpy()
        for f in range(nchk):
            assert rc[f] == counts[f] and np.array_equal(rd[f, :rc[f]], hd[f, :rc[f]]), "GPU/CPU parity broke in bench (descriptors)"
            assert np.array_equal(np.ascontiguousarray(hk[f, :rc[f]]).view(np.uint8).reshape(-1),
                                  np.ascontiguousarray(rk[f, :rc[f]]).view(np.uint8).reshape(-1)), "GPU/CPU parity broke in bench (keypoints)"
        npchk = min(3, npairs, nchk - 1)
        for p in range(npchk):
            w = port.knn2(rd[p, :rc[p]], rd[p + 1, :rc[p + 1]], TH_LOW, RATIO)
            got = [t[p, :rc[p]].cpu().numpy() for t in (blk.d_midx, blk.d_md1, blk.d_md2, blk.d_mok)]
            assert all(np.array_equal(a, b) for a, b in zip(got, w)), "GPU/CPU parity broke in bench (matches)"
        parity = "frames 0..%d and pairs 0..%d of the timed outputs equal the CPU oracle bit for bit" % (nchk - 1, max(npchk - 1, 0))

    # ---- stage shares + Hamming rate (events inside the library / around the pairs kernel) ------------------
    stage_ms = ex.profile_stages(blk.d_imgs.data_ptr(), F, W_IMG, H_IMG, W_IMG, W_IMG * H_IMG, blk.d_kps.data_ptr(), blk.d_desc.data_ptr(), cap,
                                 blk.d_counts.data_ptr(), stream)
    names = ["pyramid", "fast", "quadtree", "blur", "orient_desc"]
    ab = algorithmic_bytes()
    hbm_peak, peak_src = 6650.0, "fallback"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, peak_src = float(mp["hbm_gbs"]), "measured"
    except Exception:
        pass
    stages = {n: {"ms_per_step": float(stage_ms[i]), "GBps": ab[n] * F / (float(stage_ms[i]) * 1e6) if stage_ms[i] > 0 else None}
              for i, n in enumerate(names)}
    if npairs > 0:
        k0 = torch.cuda.Event(enable_timing=True); k1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        k0.record()
        for _ in range(3):
            api.knn2_pairs_device(blk.d_desc.data_ptr(), blk.d_counts.data_ptr(), cap, blk.d_qf.data_ptr(), blk.d_tf.data_ptr(), npairs, TH_LOW, RATIO,
                                  blk.d_midx.data_ptr(), blk.d_md1.data_ptr(), blk.d_md2.data_ptr(), blk.d_mok.data_ptr(), stream)
        k1.record()
        torch.cuda.synchronize()
        pair_ms = k0.elapsed_time(k1) / 3
        pair_matches = float((counts[:npairs].astype(np.int64) * counts[1:npairs + 1].astype(np.int64)).sum())
        stages["hamming_pairs"] = {"ms_per_step": pair_ms, "gmatch_per_s": pair_matches / (pair_ms * 1e6)}
    else:
        stages["hamming_pairs"] = {"ms_per_step": 0.0, "gmatch_per_s": 0.0}
    # integer roofline of the Hamming kernel: POPC is the scarce pipe; the carry-save distance needs 5 POPC per pair
    popc_peak, popc_src = 148 * 16 * 1.965e9, "nominal 16 POPC/clk/SM"
    try:
        ip = json.load(open(os.path.join(ROOT, "profiles", "measured_int_peaks.json")))
        popc_peak, popc_src = float(ip["popc_per_s"]), "measured (tools/int_peak.cu, profiles/measured_int_peaks.json)"
    except Exception:
        pass
    gm = stages["hamming_pairs"]["gmatch_per_s"]
    hamming_roofline = {"kernel": "knn2_pairs_kernel", "bound": "int-popc", "achieved": gm * 5e9, "peak": popc_peak, "unit": "POPC/s",
                        "frac": gm * 5e9 / popc_peak, "peak_source": popc_src,
                        "note": "5 POPC per 256-bit pair (carry-save tree); the plain 8-POPC form would be bound at %.0f Gmatch/s"
                                % (popc_peak / 8e9)}
    dom = max(names, key=lambda n: stages[n]["ms_per_step"])
    chunk = int(os.environ.get("ORBX_CHUNK", "512"))      # frames per launch of the device-resident path (orb_capi.cu kDefaultChunk)
    nchunks = (F + chunk - 1) // chunk
    frames_per_launch = F / nchunks
    dom_ms_launch = stages[dom]["ms_per_step"] / nchunks
    achieved = ab[dom] * frames_per_launch / (dom_ms_launch * 1e6)
    traffic, traffic_src = None, None
    for tf in ("r2_traffic.json", "r1_traffic.json"):
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture, per launch
            tr = json.load(open(os.path.join(ROOT, "profiles", tf)))
            knames = {"fast": ("fast_warp_kernel", "fast_kernel"), "blur": ("blur_tma_kernel", "blur_walk_kernel", "blur_kernel"),
                      "orient_desc": ("orient_desc_fused_kernel", "orient_desc_tma_kernel", "orient_desc_kernel"), "quadtree": ("octree_kernel",),
                      "pyramid": ("resize_tma_kernel", "resize_walk_kernel", "resize4_kernel")}[dom]
            kname = [k for k in knames if k in tr["kernels"]][0]
            traffic = tr["kernels"][kname]["dram_bytes_per_frame"] * frames_per_launch
            traffic_src = "profiles/" + tf
            break
        except Exception:
            pass
    per_gpu_fps = value / world
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "note": "algorithmic bytes/launch = %d B/frame x %.0f frames/launch; whole-frame figure %.2f MB/frame -> %.1f GB/s per GPU (frac %.4f)"
                        % (ab[dom], frames_per_launch, ab["frame_total"] / 1e6, ab["frame_total"] * per_gpu_fps / 1e9,
                           ab["frame_total"] * per_gpu_fps / 1e9 / hbm_peak)}

    # The binding resource of the image stages is the instruction issue rate, not HBM (ncu: DRAM 6-38 % of peak, issue slots 76-90 %
    # busy).  For the dominant stage the executed warp instructions per frame come from the committed ncu capture
    # (profiles/r2_fast_lines.md: 827,639,446 per 512 frames); achieved = that count over the live stage time, peak = 4 schedulers per
    # SM x SMs x the SM clock sampled during the run.
    if dom == "fast":
        try:
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            mhz = float(clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0)
            wi_frame = 827639446 / 512.0
            ach = wi_frame * F / (stages["fast"]["ms_per_step"] * 1e-3)     # F = frames of this rank (the stage times are this rank's)
            peak_issue = 4.0 * sms * mhz * 1e6
            roofline["issue_slots"] = {"bound": "warp-instruction issue", "warp_instructions_per_frame": wi_frame, "achieved": ach,
                                       "peak": peak_issue, "unit": "warp-instr/s", "frac": ach / peak_issue,
                                       "source": "profiles/r2_fast_lines.md (ncu smsp__inst_executed.sum) / live stage time"}
        except Exception:       # noqa: BLE001
            pass

    # ---- what bounds e2e: the pinned host->device copy rate of the same input, ALL ranks copying at the same time -------------
    link = None
    try:
        c0 = torch.cuda.Event(enable_timing=True); c1 = torch.cuda.Event(enable_timing=True)
        blk.d_imgs.copy_(blk.h_imgs, non_blocking=True)
        barrier()
        c0.record()
        for _ in range(3):
            blk.d_imgs.copy_(blk.h_imgs, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        mine_ms = c0.elapsed_time(c1)
        gbs = 3 * h2d / (mine_ms * 1e6)
        agg = gbs
        if world > 1:
            tb = torch.tensor([3.0 * h2d], dtype=torch.float64, device=dev)
            dist.all_reduce(tb)
            agg = float(tb.item()) / (max_over_ranks(mine_ms) * 1e6)
        e2e_in = e2e_value * W_IMG * H_IMG / 1e9
        link = {"h2d_copy_gbs_this_rank": gbs, "aggregate_h2d_gbs": agg, "e2e_input_gbs": e2e_in, "e2e_frac_of_aggregate": e2e_in / agg,
                "simultaneous_ranks": world, "numa_node": numa}
        # the copies of one step and nothing else: the input up on one stream while the step's results (counts, keypoints,
        # descriptors, matches) come down on another, on ALL ranks at once -- no e2e step can be shorter than this
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        H = blk.host
        pairs_rows = max(F - 1, 1)

        def copies_only():
            with torch.cuda.stream(s_up):
                blk.d_imgs.copy_(blk.h_imgs, non_blocking=True)
            with torch.cuda.stream(s_dn):
                H["counts"].copy_(blk.d_counts, non_blocking=True); H["kps"].copy_(blk.d_kps, non_blocking=True)
                H["desc"].copy_(blk.d_desc, non_blocking=True)
                for hk, dt in (("midx", blk.d_midx), ("md1", blk.d_md1), ("md2", blk.d_md2), ("mok", blk.d_mok)):
                    H[hk][:dt.shape[0]].copy_(dt[:pairs_rows], non_blocking=True)
        copies_only()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            copies_only()
        torch.cuda.synchronize()
        tc = max_over_ranks(time.perf_counter() - t0) / 3
        link["copies_only_ms_per_step"] = tc * 1e3
        link["copies_only_frames_per_s"] = job_frames / tc
        link["e2e_frac_of_copies_only"] = e2e_value / (job_frames / tc)
    except Exception as exc:                          # noqa: BLE001
        link = {"error": str(exc)[:120]}

    # ---- A/B of the opt-in Hamming variant (integer tensor pipe, identical results): the same resident step with it switched on --------
    variants = None
    if not args.skip_variants and npairs > 0:
        try:
            prev = api.set_hamming_variant(1)
            for _ in range(3):
                blk.step_resident(stream)
            v0 = torch.cuda.Event(enable_timing=True); v1 = torch.cuda.Event(enable_timing=True)
            barrier()
            v0.record()
            for _ in range(steps):
                blk.step_resident(stream)
            v1.record()
            barrier()
            vms = max_over_ranks(v0.elapsed_time(v1))
            vchk = blk.checksum()
            if world > 1:
                dist.all_reduce(vchk)
            h0 = torch.cuda.Event(enable_timing=True); h1 = torch.cuda.Event(enable_timing=True)
            h0.record()
            for _ in range(3):
                api.knn2_pairs_device(blk.d_desc.data_ptr(), blk.d_counts.data_ptr(), cap, blk.d_qf.data_ptr(), blk.d_tf.data_ptr(), npairs, TH_LOW, RATIO,
                                      blk.d_midx.data_ptr(), blk.d_md1.data_ptr(), blk.d_md2.data_ptr(), blk.d_mok.data_ptr(), stream)
            h1.record()
            torch.cuda.synchronize()
            hms = h0.elapsed_time(h1) / 3
            api.set_hamming_variant(prev)
            variants = {"hamming_mma": {"value": job_frames * steps / (vms / 1000.0), "ms_per_step": vms / steps, "hamming_pairs_ms": hms,
                                        "gmatch_per_s": stages["hamming_pairs"]["gmatch_per_s"] * stages["hamming_pairs"]["ms_per_step"] / hms,
                                        "same_results": [int(x) for x in vchk.cpu().tolist()] == chk,
                                        "what": "hamm_set_variant(1): descriptor bits as +-1 bytes on the integer tensor pipe (mma.sync IMMA.16832), "
                                                "opt-in, bit-identical; the headline `value` uses the default POPC kernels"}}
        except Exception as exc:                      # noqa: BLE001
            variants = {"error": str(exc)[:160]}
            api.set_hamming_variant(0)

    # ---- the other scaling mode beside it (N > 1): 4096 frames per GPU, every rank its own sequence ------------------------
    other = None
    if world > 1 and not args.skip_other:
        del blk
        torch.cuda.empty_cache()
        if args.scaling == "strong":
            b2 = Block(torch, ex, dev, synth, rank, 0, Ftot, Ftot, Ftot - 1); jf = world * Ftot; nm = "weak"
        else:
            lo, hi, hi_ext, p_hi = sharded.strong_block(Ftot, rank, world)
            b2 = Block(torch, ex, dev, synth, 0, lo, hi, hi_ext, p_hi); jf = Ftot; nm = "strong"
        m2 = measure(b2, jf, False)
        other = {"scaling": nm, "value": m2["value"], "ms_per_step": m2["ms_per_step"], "frames_per_gpu": b2.F,
                 "e2e": {k: m2["e2e"][k] for k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step")}}
        blk = b2

    # ---- map-scale sharded Hamming top-2 (BASELINE config 5 shape: 1k queries vs 2M rows per GPU) ---------
    hamming_map = None
    if not args.skip_map:
        hamming_map = bench_hamming_map(args, torch, dist, sharded, dev, rank, world, local_rank, barrier, max_over_ranks)

    # ---- single-frame latency through orbx_extract (the reference's actual call pattern: one frame per call) ------
    single = None
    if rank == 0 and not args.skip_single:
        import ctypes as C
        one = np.ascontiguousarray(blk.h_imgs.numpy()[0])
        skp = np.zeros(cap, api.KP_DTYPE); sde = np.zeros((cap, 32), np.uint8); n1 = C.c_int(0)
        L = vo.lib()

        def call():
            rc = L.orbx_extract(ex._h, C.c_void_p(one.ctypes.data), W_IMG, H_IMG, W_IMG, C.c_void_p(skp.ctypes.data),
                                C.c_void_p(sde.ctypes.data), cap, C.byref(n1))
            assert rc == 0
        for _ in range(5):
            call()
        t0 = time.perf_counter()
        for _ in range(50):
            call()
        single = {"ms_per_frame": (time.perf_counter() - t0) / 50 * 1e3, "api": "orbx_extract (host in, host out, synchronous)",
                  "keypoints": int(n1.value)}

    # ---- the other BASELINE configs as extra keys (rank 0, N = 1 only; a few seconds) ---------------------------
    extra_cfg = {}
    if rank == 0 and world == 1 and not args.skip_configs:
        try:
            extra_cfg = bench_other_configs(torch, vo, synth, hbm_peak)
        except Exception as exc:                      # noqa: BLE001
            extra_cfg = {"configs_error": str(exc)[:200]}

    # ---- CPU baseline beside it (rank 0, N=1 only) ---------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cores = os.cpu_count() or 1
        sample = min(F, max(cores * 64, 64))
        v, kind, t, extras = cpu_arm(blk.h_imgs.numpy()[:sample], cores)
        cpu_baseline = dict({"value": v, "unit": "frames/s", "cores": cores, "kind": kind,
                             "sample": "first %d frames of the batch: reference ORBextractor.cpp (oracle/_ref, -O3) one instance per "
                                       "core + the reference's matcher loop threaded over queries; %.1f s of wall time" % (sample, t)}, **extras)

    if rank == 0:
        accepted = checksum["accepted_matches"]
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": warm,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u8",
                "data": "synthetic",
                "config": {"workload": ("configs[1]: ONE batch of %d synthetic 640x480 frames per step, block-partitioned over %d GPU(s) with the "
                                        "boundary frame replicated (strong scaling)" % (Ftot, world) if args.scaling == "strong" else
                                        "configs[1]: batch of %d synthetic 640x480 frames per GPU per step (weak scaling)" % Ftot) +
                                       ", ORB extract (1000 features, 8 levels, 1.2, FAST 20/7) + frame-to-frame Hamming kNN (k=2, TH_LOW 50, "
                                       "ratio 0.7); frames sharded, no collective",
                           "frames_per_step": job_frames, "frames_on_rank0": F,
                           "l2": "inputs (%.2f GB on rank 0) larger than L2" % (F * W_IMG * H_IMG / 1e9),
                           "distinct_frames": min(Ftot, 1024),
                           "input_note": "the synthetic pan wraps every 1024 frames: frames i and i+1024 are identical (no kernel caches by content)",
                           "chunk_frames": chunk, "resident_lanes": int(os.environ.get("ORBX_LANES", "2")),
                           "chunk_frames_host_pipeline": int(os.environ.get("ORBX_CHUNK_HOST", os.environ.get("ORBX_CHUNK", "128"))),
                           "mean_keypoints": checksum["keypoints"] / max(job_frames, 1),
                           "accepted_matches_per_pair": accepted / max(owned_pairs, 1), "checksum": checksum,
                           "parity_spot_check": parity},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "bytes_note": "per rank (rank 0)" if world > 1 else "whole job",
                        "api": "orbx_extract_match_batch (host C ABI, pinned host buffers, copies inside the timed region)", "link": link},
                "gpu_launches": int(launches),
                "roofline": roofline, "stages": stages, "hamming_gmatch_per_s": stages["hamming_pairs"]["gmatch_per_s"],
                "hamming_roofline": hamming_roofline, "hamming_map": hamming_map, "single_frame": single, "cpu_baseline": cpu_baseline}
        if other is not None:
            line["other_scaling"] = other
        if variants is not None:
            line["variants"] = variants
        line.update(extra_cfg)
        emit(line)
    ex.close()


def bench_hamming_map(args, torch, dist, sharded, dev, rank, world, local_rank, barrier, max_over_ranks):
    """BASELINE config 5 shape (1000 queries x map_rows rows per GPU).  The map holds planted near-duplicates of the queries (the
    ratio test fires) and exact duplicates straddling every shard boundary (index ties across shards); a 1k x 64k sub-problem of
    the same construction is compared with the CPU oracle through BOTH exchange paths on every rank."""
    Q, Ml = 1000, args.map_rows
    gq = torch.Generator(device=dev); gq.manual_seed(99)
    d_q = torch.randint(0, 256, (Q, 32), dtype=torch.uint8, device=dev, generator=gq)

    def make_shard(rows, seed):
        g = torch.Generator(device=dev); g.manual_seed(seed + rank)
        t = torch.randint(0, 256, (rows, 32), dtype=torch.uint8, device=dev, generator=g)
        if rows >= 4 * Q:
            pos = torch.randperm(rows - 2, device=dev, generator=g)[:Q // 4] + 1          # planted near-duplicates: 0..40 flipped bits
            src = torch.randint(0, Q, (Q // 4,), device=dev, generator=g)
            nflip = torch.randint(0, 41, (Q // 4,), device=dev, generator=g)
            bits = torch.rand((Q // 4, 256), device=dev, generator=g).argsort(dim=1)
            flip = (torch.arange(256, device=dev)[None, :] < nflip[:, None])
            mask = torch.zeros((Q // 4, 256), dtype=torch.uint8, device=dev)
            mask.scatter_(1, bits, flip.to(torch.uint8))
            w = (2 ** torch.arange(8, device=dev)).to(torch.int32)
            mbytes = (mask.view(Q // 4, 32, 8).to(torch.int32) * w).sum(dim=2).to(torch.uint8)
            t[pos] = d_q[src] ^ mbytes
        # exact duplicates on both sides of the shard boundary: query r ties at distance 0 between shard r-1 and shard r
        if rows >= 2:
            if rank > 0:
                t[0] = d_q[rank]
            if rank + 1 < world:
                t[rows - 1] = d_q[rank + 1]
        return t

    xchg, merge_kind = None, "single shard"
    if world > 1:
        merge_kind = "nccl all_gather + merge kernel"
        if os.environ.get("ORBX_PEER_MERGE", "1") != "0":
            try:
                xchg = sharded.PeerExchange(dist, local_rank, Q)      # agrees on success across ranks itself
                merge_kind = "fused: P2P stores into peers' exchange buffers (CUDA IPC over NVLink) + flag wait"
            except Exception as e:      # noqa: BLE001
                sys.stderr.write("peer exchange unavailable, using NCCL all-gather: %r\n" % (e,))
                xchg = None

    def run_map(d_t, lo, use_peer):
        if use_peer:
            return sharded.sharded_knn2_peer(d_q, d_t, lo, TH_LOW, RATIO, xchg, check=False)
        return sharded.sharded_knn2_cuda(d_q, d_t, lo, TH_LOW, RATIO, dist if world > 1 else None)

    # -- oracle-checked sub-problem: 1k x 64k rows over all ranks ---------------------------------------------
    sub = 65536 // world
    d_ts = make_shard(sub, 4321)
    equals_oracle, oracle_detail = None, None
    try:
        outs = {}
        if xchg is not None:
            outs["fused"] = [x.cpu().numpy() for x in run_map(d_ts, rank * sub, True)]
        outs["nccl" if world > 1 else "single"] = [x.cpu().numpy() for x in run_map(d_ts, rank * sub, False)]
        if world > 1:
            allt = torch.empty((world, sub, 32), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allt, d_ts)
            full = allt.view(-1, 32).cpu().numpy()
        else:
            full = d_ts.cpu().numpy()
        import oracle
        want = oracle.Port().knn2(d_q.cpu().numpy(), full, TH_LOW, RATIO, nthreads=max(1, (os.cpu_count() or 1) // world))
        good = all(all(np.array_equal(a, b) for a, b in zip(o, want)) for o in outs.values())
        flag = torch.tensor([1 if good else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        equals_oracle = bool(flag.item())
        oracle_detail = {"paths": sorted(outs), "queries": Q, "rows": sub * world, "accepted": int(want[3].sum()),
                         "cross_shard_ties": int(((want[1] == want[2]) & (want[1] == 0)).sum())}
    except Exception as e:      # noqa: BLE001
        equals_oracle, oracle_detail = None, {"error": str(e)[:160]}
    del d_ts

    # -- the timed map ---------------------------------------------------------------------------------------------
    d_t = make_shard(Ml, 1234)
    lo = rank * Ml

    def time_map(use_peer):
        for _ in range(2):
            run_map(d_t, lo, use_peer)
        m0 = torch.cuda.Event(enable_timing=True); m1 = torch.cuda.Event(enable_timing=True)
        barrier()
        m0.record()
        reps = 5
        for _ in range(reps):
            out = run_map(d_t, lo, use_peer)
        m1.record()
        barrier()
        return max_over_ranks(m0.elapsed_time(m1)) / reps, out

    mms, out = time_map(xchg is not None)
    res = {"gmatch_per_s": world * Q * Ml / (mms * 1e6), "ms": mms, "queries": Q, "rows_per_gpu": Ml, "merge": merge_kind,
           "accepted": int(out[3].sum().item()), "equals_oracle": equals_oracle, "equals_oracle_detail": oracle_detail}
    if not args.skip_variants:
        try:
            from vo_slam_test_b200 import api
            prev = api.set_hamming_variant(1)
            vms, vout = time_map(xchg is not None)
            api.set_hamming_variant(prev)
            res["variant_hamming_mma"] = {"ms": vms, "gmatch_per_s": world * Q * Ml / (vms * 1e6),
                                          "same_results": all(bool(torch.equal(a, b)) for a, b in zip(out, vout))}
        except Exception as exc:                      # noqa: BLE001
            res["variant_hamming_mma"] = {"error": str(exc)[:160]}
    if xchg is not None:
        nms, ref = time_map(False)
        res["nccl_all_gather_ms"] = nms
        res["fused_equals_nccl_path"] = all(bool(torch.equal(a, b)) for a, b in zip(out, ref))
        xchg.close()
    return res


def bench_tracking_frame(vo, synth):
    """The reference's real per-frame call pattern (frame.cpp:22-32, then visualOdometry.cpp:240,265,354): ONE frame through
    Frame::Frame (extract + undistort + depth + grid) and three searches against it, timed on the host around the C-ABI calls
    (arguments marshalled once, outside the timed loop).  `handle`: orbx_frame_create + the *_h searches (frame resident);
    `host_arrays`: orbx_extract + orbx_frame_finish + the host-array searches (every call re-uploads the frame)."""
    import ctypes as C
    from vo_slam_test_b200 import api
    L = vo.lib()
    ex = vo.ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH)
    cam = vo.camera(517.3, 516.5, 318.6, 255.3, [0.2624, -0.9531, -0.0054, 0.0026, 1.1633], 40.0, (0.0, 640.0, 0.0, 480.0))
    img = synth.make_frame(42)
    rng = np.random.default_rng(5)
    depth = rng.uniform(0.5, 8.0, (H_IMG, W_IMG)).astype(np.float32)
    fr = vo.Frame(ex, cam, img, depth)
    sf = np.asarray(ex.GetScaleFactors(), np.float32)
    n = fr.n
    frame, pts = synth.make_projection_case(fr.unkps, fr.desc, sf, 1000, seed=1)
    frame["uright"] = fr.uright
    frameL, ptsL = synth.make_projection_case(fr.unkps, fr.desc, sf, 2000, seed=2, local=True)
    src = rng.integers(0, n, 900)
    kd = synth.flip_bits(fr.desc[src], rng.integers(0, 40, 900), rng)
    A = synth.make_bow_side(kd, ((fr.unkps["angle"][src] + rng.normal(0, 8, 900)) % 360).astype(np.float32), None, 9, 1)
    B = synth.make_bow_side(fr.desc, fr.unkps["angle"], None, 9, 2)
    keep = []
    M = vo.Matcher(0.9)
    sF = M._frame_points(pts, keep)

    def a(x, dt):
        y = np.ascontiguousarray(x, dt); keep.append(y); return y.ctypes.data
    sL = api._SbpLocalPoints()
    sL.m = len(ptsL["u"]); sL.valid = a(ptsL["valid"], np.uint8); sL.u = a(ptsL["u"], np.float32); sL.v = a(ptsL["v"], np.float32)
    sL.ur = a(ptsL["ur"], np.float32); sL.level = a(ptsL["level"], np.int32); sL.view_cos = a(ptsL["view_cos"], np.float32)
    sL.desc = a(ptsL["desc"], np.uint8); sL.has_obs = a(ptsL["has_obs"], np.uint8)

    def side(d, with_desc=True):
        b = api._BowSide()
        b.n = len(d["desc"]); b.desc = a(d["desc"], np.uint8) if with_desc else None; b.angle = a(d["angle"], np.float32) if with_desc else None
        b.valid = a(d["valid"], np.uint8); b.ngroups = len(d["node_ids"]); b.node_ids = a(d["node_ids"], np.uint32)
        b.group_start = a(d["group_start"], np.int32); b.feat_idx = a(d["feat_idx"], np.int32)
        return b
    bA, bB, bBh = side(A), side(B), side(B, False)
    fv = M._frame_view(frame, keep)
    occ = a(frame["occupied0"], np.uint8)
    cap = ex.max_keypoints
    assign = np.zeros(cap, np.int32); cnt = C.c_int(0); nn = C.c_int(0)
    vp = C.c_void_p
    kps = np.zeros(cap, api.KP_DTYPE); desc = np.zeros((cap, 32), np.uint8); un = np.zeros(cap, api.KP_DTYPE)
    ur = np.zeros(cap, np.float32); dp = np.zeros(cap, np.float32); cs = np.zeros(64 * 48 + 1, np.int32); ids = np.zeros(cap, np.int32)
    cnt1 = np.zeros(1, np.int32)
    pimg, pdepth = vp(img.ctypes.data), vp(depth.ctypes.data)
    pa = vp(assign.ctypes.data)
    fr.close()
    h = vp()

    def one_handle():
        assert L.orbx_frame_create(ex._h, C.byref(cam), pimg, W_IMG, H_IMG, W_IMG, pdepth, 4 * W_IMG, C.byref(h), C.byref(nn)) == 0
        L.orbx_frame_get(h, vp(kps.ctypes.data), vp(desc.ctypes.data), vp(un.ctypes.data), vp(ur.ctypes.data), vp(dp.ctypes.data), cap)
        r1 = L.orbx_search_by_projection_frame_h(h, vp(occ), C.byref(sF), 15.0, 40.0, 0, 0, 1, pa, C.byref(cnt))
        r2 = L.orbx_search_by_bow_h(C.byref(bA), h, C.byref(bBh), 0.7, TH_LOW, 1, pa, C.byref(cnt))
        r3 = L.orbx_search_by_projection_local_h(h, vp(occ), C.byref(sL), 3.0, 0.8, pa, C.byref(cnt))
        L.orbx_frame_destroy(h)
        assert r1 == 0 and r2 == 0 and r3 == 0

    def one_host():
        assert L.orbx_extract(ex._h, pimg, W_IMG, H_IMG, W_IMG, vp(kps.ctypes.data), vp(desc.ctypes.data), cap, C.byref(nn)) == 0
        cnt1[0] = nn.value
        assert L.orbx_frame_finish(C.byref(cam), vp(kps.ctypes.data), vp(cnt1.ctypes.data), 1, cap, pdepth, W_IMG, H_IMG, 4 * W_IMG, 0,
                                   vp(un.ctypes.data), vp(ur.ctypes.data), vp(dp.ctypes.data), vp(cs.ctypes.data), vp(ids.ctypes.data), 0) == 0
        r1 = L.orbx_search_by_projection_frame(C.byref(fv), C.byref(sF), 15.0, 40.0, 0, 0, 1, pa, C.byref(cnt), 0)
        r2 = L.orbx_search_by_bow(C.byref(bA), C.byref(bB), 0, 0.7, TH_LOW, 1, pa, C.byref(cnt), 0)
        r3 = L.orbx_search_by_projection_local(C.byref(fv), C.byref(sL), 3.0, 0.8, pa, C.byref(cnt), 0)
        assert r1 == 0 and r2 == 0 and r3 == 0

    out = {}
    for name, fn in (("handle", one_handle), ("host_arrays", one_host)):
        for _ in range(10):
            fn()
        t0 = time.perf_counter()
        for _ in range(100):
            fn()
        out[name] = (time.perf_counter() - t0) / 100 * 1e3
    # the parts of the handle path
    parts = {}

    def timeit(fn, reps=100):
        for _ in range(5):
            fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e3
    assert L.orbx_frame_create(ex._h, C.byref(cam), pimg, W_IMG, H_IMG, W_IMG, pdepth, 4 * W_IMG, C.byref(h), C.byref(nn)) == 0
    parts["search_by_projection_frame_h"] = timeit(lambda: L.orbx_search_by_projection_frame_h(h, vp(occ), C.byref(sF), 15.0, 40.0, 0, 0, 1, pa, C.byref(cnt)))
    parts["search_by_bow_h"] = timeit(lambda: L.orbx_search_by_bow_h(C.byref(bA), h, C.byref(bBh), 0.7, TH_LOW, 1, pa, C.byref(cnt)))
    parts["search_by_projection_local_h"] = timeit(lambda: L.orbx_search_by_projection_local_h(h, vp(occ), C.byref(sL), 3.0, 0.8, pa, C.byref(cnt)))
    L.orbx_frame_destroy(h)

    def create_only():
        L.orbx_frame_create(ex._h, C.byref(cam), pimg, W_IMG, H_IMG, W_IMG, pdepth, 4 * W_IMG, C.byref(h), C.byref(nn)); L.orbx_frame_destroy(h)
    parts["frame_create_with_depth"] = timeit(create_only)

    def create_nodepth():
        L.orbx_frame_create(ex._h, C.byref(cam), pimg, W_IMG, H_IMG, W_IMG, None, 0, C.byref(h), C.byref(nn)); L.orbx_frame_destroy(h)
    parts["frame_create_no_depth"] = timeit(create_nodepth)
    ex.close()
    return {"tracking_frame_ms": out["handle"], "tracking_frame_ms_host_arrays": out["host_arrays"], "parts_ms": parts,
            "what": "one 640x480 frame: Frame::Frame (extract 1000 kp + undistort + findDepth from a pageable 640x480 float depth image + grid) "
                    "+ searchByProjection(Frame,Frame) 1000 pts + searchByBoW(KeyFrame,Frame) 900 feats + searchByProjection(local map) 2000 pts; "
                    "host wall clock around the C-ABI calls, 100 reps"}


def bench_other_configs(torch, vo, synth, hbm_peak):
    """BASELINE configs 3 (1920x1080 / 2000 features, 3840x2160 / 5000 features) and 4 (10k map points projected into one VGA
    frame, radius 15) as extra keys of the bench line; a few seconds in total."""
    out = {}

    def lv_sizes(W, H):
        sc = [1.0]
        for _ in range(1, NLEVELS):
            sc.append(float(np.float32(sc[-1] * float(np.float32(SCALE)))))
        return [(int(np.rint(np.float32(W) * (np.float32(1.0) / np.float32(s)))), int(np.rint(np.float32(H) * (np.float32(1.0) / np.float32(s))))) for s in sc]

    def resident(W, H, nfeat, B):
        ex = vo.ORBextractor(nfeat, SCALE, NLEVELS, INI_TH, MIN_TH)
        cap = ex.max_keypoints
        base = np.stack([synth.make_frame(100 + i, H, W) for i in range(2)])
        imgs = np.concatenate([base] * ((B + 1) // 2))[:B]
        d = torch.from_numpy(imgs).cuda()
        k = torch.empty((B, cap, 7), dtype=torch.float32, device="cuda"); de = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda")
        c = torch.zeros(B, dtype=torch.int32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        run = lambda: ex.extract_batch_device(d.data_ptr(), B, W, H, W, W * H, k.data_ptr(), de.data_ptr(), cap, c.data_ptr(), st)   # noqa: E731
        for _ in range(3):
            run()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        reps = 5
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        sz = lv_sizes(W, H)
        A = sum(w * h for w, h in sz)
        nk = float(c.float().mean().item())
        fbytes = 5 * A - sz[0][0] * sz[0][1] - sz[-1][0] * sz[-1][1] + 60 * nk
        fps = B / (ms / 1e3)
        ex.close()
        return {"frames_per_s": fps, "batch": B, "ms_per_batch": ms, "mean_keypoints": nk, "algorithmic_MB_per_frame": fbytes / 1e6,
                "achieved_GBps": fbytes * fps / 1e9, "frac": fbytes * fps / 1e9 / hbm_peak, "timing": "resident, CUDA events, 5 reps after 3 warm-ups"}
    out["config3_1080p"] = resident(1920, 1080, 2000, 128)
    out["config3_4k"] = resident(3840, 2160, 5000, 32)
    # config 4
    exv = vo.ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH)
    kps, desc = exv(synth.make_frame(42))
    sf = np.asarray(exv.GetScaleFactors(), np.float32)
    exv.close()
    res = {}
    for name, ratio, kw, fn in (("frame", 0.9, {}, "f"), ("frame_stereo", 0.9, {"stereo": True}, "f"), ("local_map", 0.8, {"local": True}, "l")):
        frame, pts = synth.make_projection_case(kps, desc, sf, 10000, seed=1, **kw)
        M = vo.Matcher(ratio)
        call = (lambda: M.searchByProjection(frame, pts, 15.0)) if fn == "f" else (lambda: M.searchByProjectionLocal(frame, pts, 3.0))
        for _ in range(3):
            r = call()
        t0 = time.perf_counter()
        for _ in range(20):
            r = call()
        dt = (time.perf_counter() - t0) / 20
        res[name] = {"ms_per_search": dt * 1e3, "points_per_s": 10000 / dt, "matches": int(r[1]) if isinstance(r, tuple) else None}
    res["api"] = "host C ABI (orbx_search_by_projection_frame / _local): host arrays in, assignment out, synchronous"
    # the same searches against the frame kept resident by orbx_frame_create (zero distortion: unKeypoints_ == keypoints_)
    try:
        exh = vo.ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH)
        cam0 = vo.camera(517.3, 516.5, 318.6, 255.3, [0.0, 0.0, 0.0, 0.0, 0.0], 40.0, (0.0, 640.0, 0.0, 480.0))
        fr = vo.Frame(exh, cam0, synth.make_frame(42))
        for name, ratio, kw, fn in (("frame_h", 0.9, {}, "f"), ("local_map_h", 0.8, {"local": True}, "l")):
            frame, pts = synth.make_projection_case(kps, desc, sf, 10000, seed=1, **kw)
            M = vo.Matcher(ratio)
            call = ((lambda: M.searchByProjectionH(fr, frame["occupied0"], pts, 15.0)) if fn == "f"
                    else (lambda: M.searchByProjectionLocalH(fr, frame["occupied0"], pts, 3.0)))
            for _ in range(3):
                r = call()
            t0 = time.perf_counter()
            for _ in range(20):
                r = call()
            dt = (time.perf_counter() - t0) / 20
            res[name] = {"ms_per_search": dt * 1e3, "points_per_s": 10000 / dt, "matches": int(r[1]),
                         "same_matches_as_host_arrays": int(r[1]) == res[name[:-2]]["matches"]}
        fr.close(); exh.close()
        res["api_h"] = "orbx_search_by_projection_frame_h / _local_h on the resident frame (orbx_frame_t)"
    except Exception as exc:                          # noqa: BLE001
        res["handle_error"] = str(exc)[:200]
    out["config4_sbp"] = res
    try:
        out["tracking_frame"] = bench_tracking_frame(vo, synth)
        out["tracking_frame_ms"] = out["tracking_frame"]["tracking_frame_ms"]
    except Exception as exc:                          # noqa: BLE001
        out["tracking_frame"] = {"error": str(exc)[:200]}
    return out


_REAL_STDOUT = None


def emit(line):
    """The one JSON line goes to the real stdout; everything else any library prints (e.g. NCCL's version banner)
    was redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode()); sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=4096, help="frames per step: of the whole job (strong) or per GPU (weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default, BASELINE configs[1] as written): ONE batch split over the GPUs with the boundary frame "
                         "replicated; weak: a batch per GPU")
    ap.add_argument("--skip-other", action="store_true", help="N > 1: do not also measure the other scaling mode")
    ap.add_argument("--skip-configs", action="store_true", help="do not add BASELINE configs 3/4 as extra keys")
    ap.add_argument("--skip-variants", action="store_true", help="do not A/B the opt-in Hamming variant")
    ap.add_argument("--map-rows", type=int, default=2097152, help="map descriptors per GPU for the sharded kNN side metric")
    ap.add_argument("--skip-map", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-single", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
