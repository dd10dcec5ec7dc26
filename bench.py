#!/usr/bin/env python
"""Benchmark of the post-processing hot path (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload db|pse|pan|ctc|crop|db_fp16|pse_api|pan_api] [--batch B] [--no-cpu] [--headline-only]

A "step" is one pass of the hot path over one batch of synthetic input. The headline workload is
BASELINE.json configs[1]: DB++ r18 post-process, batch 256 synthetic 736x1280 maps (~200 text
regions each) per GPU. The default run then also times BASELINE.json configs[2..4] (PSE batch 16, PAN batch 128,
CTC 8192 lines per GPU), the two API-faithful twins (1/4-resolution head outputs through the operators' own
up-sampling) and the fp16 DB variant for a few steps each and reports them under `other_workloads` on the same
JSON line (same measurement rules; headline keys unchanged).
Scaling is WEAK: every rank processes its own batch (the path shards by image / text line, no
collective on the data path); `value` = units all ranks processed / max-over-ranks device time.
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

H, W = 736, 1280
SEED = 20221001
DB_CFG = dict(thresh=0.3, box_thresh=0.5, max_candidates=1000, unclip_ratio=1.7, score_mode="poly", cpp_speedup=True)
PSE_CFG = dict(thresh=0, box_thresh=0.85, min_area=16, scale=1)                          # det_r50_pse.yml:56-62
PAN_CFG = dict(thresh=0, box_thresh=0.85, min_area=16, min_kernel_area=2.6, scale=1)     # det_r18_pan.yml:62-69 at full res
CTC_T, CTC_C = 80, 6623
CROP_BOXES = 200


# ----------------------------------------------------------------------------------------------
# host-side workers (fork pool; run BEFORE CUDA is initialised)
# ----------------------------------------------------------------------------------------------
def _gen_db(seed):
    from pytorchocr_b200 import synth
    return synth.db_map(seed, H, W)


def _gen_pse_scene(seed):
    from pytorchocr_b200 import synth
    return synth.pse_scene(seed, H, W)[:2]


def _gen_pan_scene(seed):
    from pytorchocr_b200 import synth
    return synth.pan_scene(seed, H, W)[:4]


_SHARED = {}    # arrays the CPU workers inherit through fork (nothing is pickled through the pool's pipes)


def _cpu_db(i):
    """The reference's CPU path for map i of the inherited batch: db_postprocess.py:43-46 (threshold, uint8 cast)
    + BoxesFromBitmap (oracle/db_oracle_fast.py: the oracle's C++-branch semantics, vectorised between the same cv2 /
    Clipper calls so that Python interpreter overhead is not billed to the reference)."""
    import cv2
    cv2.setNumThreads(1)
    from oracle import db_oracle_fast
    m = _SHARED["db_maps"][i, 0]
    boxes = db_oracle_fast.boxes_from_bitmap(m, (m > DB_CFG["thresh"]).astype(np.uint8), DB_CFG["box_thresh"],
                                             DB_CFG["unclip_ratio"], W, H)
    return len(boxes)


def _cpu_pse(seed):
    import cv2
    cv2.setNumThreads(1)
    from oracle.pse_oracle import PSEPostProcessOracle
    from pytorchocr_b200 import synth
    m = synth.pse_maps(seed, H, W)
    t0 = time.perf_counter()
    r = PSEPostProcessOracle(maps_at_processing_res=True, **PSE_CFG)({"maps": m[None]}, [[H, W, 1.0, 1.0]])
    return time.perf_counter() - t0, len(r[0]["points"])


def _cpu_pan(seed):
    import cv2
    cv2.setNumThreads(1)
    from oracle.pan_oracle import PANPostProcessOracle
    from pytorchocr_b200 import synth
    m = synth.pan_maps(seed, H, W)
    t0 = time.perf_counter()
    r = PANPostProcessOracle(maps_at_processing_res=True, **PAN_CFG)({"maps": m[None]}, [[H, W, 1.0, 1.0]])
    return time.perf_counter() - t0, len(r[0]["points"])


def _cpu_ctc(args):
    seed, lines, dict_path = args
    from oracle.ctc_oracle import CTCLabelDecodeNumpy
    from pytorchocr_b200 import synth
    probs, _ = synth.ctc_probs_numpy(seed, CTC_T, lines, CTC_C)
    op = CTCLabelDecodeNumpy(dict_path)
    x = np.ascontiguousarray(probs.transpose(1, 0, 2))   # numpy input is [B,T,C] (rec_postprocess.py:80-82)
    t0 = time.perf_counter()
    r = op(x)
    return time.perf_counter() - t0, len(r)


def _gen_page(seed):
    from pytorchocr_b200 import synth
    return synth.page_image(seed, H, W), synth.page_boxes(seed + 100000, n=CROP_BOXES, H=H, W=W)


def _cpu_crop(page):
    import cv2
    cv2.setNumThreads(1)
    from oracle import crop_oracle as co
    img, boxes = page
    t0 = time.perf_counter()
    n = 0
    for b in co.sort_boxes(boxes):          # run_ocr.py:181-191: sort, crop, rot90 rule
        n += co.crop_for_rec(img, b).size
    return time.perf_counter() - t0, n


def _timed_pool(pool, fn, items, cores):
    """Runs fn over items on the pool; returns (outputs, seconds of wall time)."""
    pool.map(fn, items[:cores])  # warm the workers (imports, page-in)
    t0 = time.perf_counter()
    out = pool.map(fn, items, chunksize=1)
    dt = time.perf_counter() - t0
    return out, dt


def _cpu_db_batch(maps, cores, passes=1):
    """Times the CPU path over `maps` ([B,1,H,W] float32): a fresh fork pool inherits the array; >= 8 tasks per
    worker at B = 256 on 32 cores. Returns (maps per second, boxes found, seconds)."""
    import multiprocessing as mp
    _SHARED["db_maps"] = maps
    idx = list(range(len(maps)))
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_db, idx[:cores])
        t0 = time.perf_counter()
        for _ in range(passes):
            out = pool.map(_cpu_db, idx, chunksize=1)
        dt = time.perf_counter() - t0
    return len(maps) * passes / dt, int(sum(out)), dt


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler(object):
    """SM clock and throttle reasons of one GPU, sampled DURING the timed regions. Reads NVML in-process
    (nvidia_ml_py; ~50 us per sample, no fork): forking `nvidia-smi` every 100 ms from every rank inside a 15 ms
    event-timed window was the source of the per-rank jitter of the round-1 scaling run. Falls back to
    `nvidia-smi` when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples = []      # (sm_mhz, sm_max_mhz, set of reasons)
        self._stop = threading.Event()
        self._t = None
        self.how = "nvml"
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[gpu_index]) if visible and visible.split(",")[gpu_index].isdigit() else gpu_index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nv = pynvml
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None
            self.how = "nvidia-smi"

    def _sample_nvml(self):
        nv = self._nv
        sm = float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        names = set()
        for nm, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
                        ("hw_thermal_slowdown", 0x40)):
            if r & bit:
                names.add(nm)
        self.samples.append((sm, self._max, names))

    def _sample_smi(self):
        out = subprocess.check_output(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits"], timeout=5).decode().strip()
        f = [x.strip() for x in out.split(",")]
        names = {nm for nm, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9])
                 if v.lower().startswith("active")}
        self.samples.append((float(f[1]), float(f[2]), names))

    def _run(self):
        while not self._stop.is_set():
            try:
                (self._sample_nvml if self._nv else self._sample_smi)()
            except Exception:
                pass
            self._stop.wait(0.02 if self._nv else 0.25)

    def start(self):
        self._stop.clear()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        reasons = set()
        for smp in self.samples:
            reasons |= smp[2]
        return {"sm_mhz": float(np.median([x[0] for x in self.samples])), "sm_max_mhz": float(max(x[1] for x in self.samples)),
                "reasons": sorted(reasons), "samples": len(self.samples), "how": self.how}


# ----------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------
class Workload(object):
    name = None
    unit = "images/s"
    metric = "post-process images/sec @736x1280"
    default_batch = 256
    dtype = "f32"
    stream_kernel = None     # the kernel that streams the operator input (phase 0 of the profile)

    def __init__(self, args, rank, world):
        self.args, self.rank, self.world = args, rank, world
        if os.environ.get("OCRPP_BENCH_DATA_RANK"):   # development aid: the synthetic pages of another rank on one GPU
            self.rank = int(os.environ["OCRPP_BENCH_DATA_RANK"])
        self.batch = args.batch or self.default_batch
        self.cores = max(1, (os.cpu_count() or 1) // max(1, world))

    # per-unit algorithmic bytes: the operator input read once
    def alg_bytes_per_unit(self):
        raise NotImplementedError

    def needs_flush(self):
        return False

    def release(self):
        for k in list(vars(self)):
            if k not in ("args", "rank", "world", "batch", "cores"):
                delattr(self, k)
        import torch
        torch.cuda.empty_cache()


class DetWorkload(Workload):
    """Shared plumbing of the three detection workloads."""
    channels = 1
    op_name = None
    cfg = None
    in_h, in_w = H, W          # resolution of the operator input (the API-faithful twins feed 1/4-resolution maps)
    elem_bytes = 4

    def alg_bytes_per_unit(self):
        return self.channels * self.in_h * self.in_w * self.elem_bytes

    def device_prepare(self, dev):
        import torch
        from pytorchocr_b200.postprocess import build_post_process
        self.torch, self.dev = torch, dev
        self.op = build_post_process(dict(self.cfg, name=self.op_name, cuda_speedup=True, **self.extra_cfg()))
        self.dev_maps = self.make_device_maps(dev)
        self.host_maps = torch.empty(self.dev_maps.shape, dtype=self.dev_maps.dtype, pin_memory=True)
        self.host_maps.copy_(self.dev_maps)
        self.shape_list = np.array([[H, W, 1.0, 1.0]] * self.batch, np.float64)
        res = self.op({"maps": self.dev_maps}, self.shape_list)   # allocates/caches buffers, settles capacities
        self.n_boxes = int(sum(len(r["points"]) for r in res))
        self.buf = next(iter(self.op._cache.values()))
        self.key = next(iter(self.op._cache))

    def extra_cfg(self):
        return {}

    # ---- results of step k travel to pinned host memory while the kernels of step k+1 run: two output buffers, a
    #      copy stream, and a join before the end event of the timed region (device_join). With a single buffer the
    #      2 MB result copy sits between two steps of the same stream - 40 us at 50 GB/s, a multiple of that on a
    #      rank whose PCIe path is shared (the slow ranks of the 4- and 8-GPU runs) ----
    def out_pair(self, stream):
        torch = self.torch
        if not hasattr(self, "_outs"):
            self._outs = [(self.buf["out_dev"], self.buf["out_host"]),
                          (torch.empty_like(self.buf["out_dev"]), torch.empty_like(self.buf["out_host"]).pin_memory())]
            self._copy_stream = torch.cuda.Stream()
            self._done = [torch.cuda.Event(), torch.cuda.Event()]
            self._copied = [None, None]
            self._k = 0
        i = self._k & 1
        self._k += 1
        if self._copied[i] is not None:
            stream.wait_event(self._copied[i])     # step k-2 has left this buffer
        return i, self._outs[i][0], self._outs[i][1]

    def out_copy(self, i, stream):
        torch = self.torch
        self._done[i].record(stream)
        self._copy_stream.wait_event(self._done[i])
        with torch.cuda.stream(self._copy_stream):
            self._outs[i][1].copy_(self._outs[i][0], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._copied[i] = ev

    def device_join(self, stream):
        """every result copy issued so far is ordered before whatever is recorded on `stream` next"""
        for ev in getattr(self, "_copied", []):
            if ev is not None:
                stream.wait_event(ev)

    def e2e_step(self):
        d = self.host_maps.to(self.dev, non_blocking=True)       # H2D of this step's maps from pinned memory
        return self.op({"maps": d}, self.shape_list)             # kernels + D2H of boxes/counts + host assembly

    def h2d_bytes(self):
        return int(self.host_maps.numel() * self.host_maps.element_size() + self.shape_list.size * 8)


    def d2h_bytes(self):
        return int(self.buf["out_host"].numel())

    def public_config(self):
        """The workload as both arms name it (identical dicts in `config` of our line and of --impl reference)."""
        return {"batch_per_gpu": self.batch, "H": H, "W": W, "l2": self.l2_note()}

    def detail(self):
        return {"input": list(self.dev_maps.shape[1:]) + [str(self.dev_maps.dtype)], "boxes_per_step_rank0": self.n_boxes,
                "timed_region": "device maps -> boxes/scores/counts in pinned host memory; the result copy of step k runs "
                                "on a copy stream beside the kernels of step k+1 (two output buffers) and is joined "
                                "before the end event"}

    def l2_note(self):
        mb = self.batch * self.alg_bytes_per_unit() / 1e6
        if mb > 2 * 126:
            return "inputs (%.0f MB per GPU) larger than the 126 MB L2; no flush needed" % mb
        return "inputs are %.0f MB per GPU: a 512 MB buffer is overwritten between timed steps to flush the L2" % mb

    def needs_flush(self):
        return self.batch * self.alg_bytes_per_unit() <= 2 * 126e6


class DbWorkload(DetWorkload):
    name = "db"
    default_batch = 256
    op_name = "DBPostProcess"
    cfg = DB_CFG
    stream_kernel = "db_scan4_kernel"
    workload = ("DB++ r18 post-process, batch 256 synthetic 736x1280 maps with ~200 text regions each per GPU "
                "(BASELINE.json configs[1])")
    cpu_note = ("oracle/db_oracle_fast.py = the oracle's db_postprocess.cpp semantics (cv2-python 4.13 + the reference's "
                "own compiled Clipper) with the arithmetic between the cv2 calls vectorised per image; the C++ module "
                "itself needs OpenCV C++ and cannot be built here")
    map_dtype = "float32"

    def host_prepare(self, pool, want_cpu):
        maps = pool.map(_gen_db, [SEED + self.rank * self.batch + i for i in range(self.batch)], chunksize=4)
        self.maps = np.stack(maps)[:, None]
        if not want_cpu:
            return None
        rate, nboxes, dt = _cpu_db_batch(self.maps, self.cores)
        return {"value": rate, "unit": self.unit, "cores": self.cores, "kind": "port",
                "sample": "all %d maps of this step (the ones the GPU arm processes; %d boxes) in %.1f s, fork pool of %d "
                          "workers inheriting the array, cv2.setNumThreads(1); %s"
                          % (len(self.maps), nboxes, dt, self.cores, self.cpu_note)}

    def make_device_maps(self, dev):
        t = self.torch.from_numpy(self.maps).pin_memory().to(dev)
        return t.half() if self.map_dtype == "float16" else t

    def e2e_step(self):
        # the operator takes the pinned HOST maps (as the reference's operator takes `.cpu().numpy()` maps) and
        # does the H2D itself, chunked and overlapped with the kernels of the previous chunk
        return self.op({"maps": self.host_maps}, self.shape_list)

    def device_step(self, L, stream):
        from pytorchocr_b200 import _lib
        i, out_dev, _ = self.out_pair(stream)
        # The call is ~25 driver calls (memset, fork/join of the sub-batch streams, 8 launches) for 0.41 ms of GPU
        # work: on a host whose cores are shared (the slow ranks of the 4- and 8-GPU runs) the enqueue itself took
        # longer than the kernels. The step is therefore captured ONCE per output buffer into a CUDA graph (stream
        # capture follows the library's fork/join onto its auxiliary streams) and replayed; the per-kernel profiling
        # pass and any failure to capture use direct launches.
        if getattr(self, "graph_ok", True) and not getattr(self, "_graph_failed", False):
            g = self._graphs.get(i) if hasattr(self, "_graphs") else None
            if g is None:
                g = self._capture(L, i, out_dev)
            if g is not None:
                g[0].replay()
                self.replayed_launches = getattr(self, "replayed_launches", 0) + g[1]
                self.out_copy(i, stream)
                return
        self._launch(L, stream, out_dev)
        self.out_copy(i, stream)

    def _launch(self, L, stream, out_dev):
        from pytorchocr_b200 import _lib
        buf, m = self.buf, self.dev_maps
        o_box, o_sc, o_cnt, o_st = buf["offs"]
        base = out_dev.data_ptr()
        _lib.check(L.ocrpp_db_postprocess(
            m.data_ptr(), _lib.F16 if self.map_dtype == "float16" else _lib.F32, self.batch, H, W, m.stride(0),
            m.stride(2), buf["wh_dev"].data_ptr(),
            DB_CFG["thresh"], DB_CFG["box_thresh"], DB_CFG["unclip_ratio"], self.key[5], self.key[4], 0, 0,
            base + o_box, base + o_sc, base + o_cnt, base + o_st, None, None,
            buf["ws"].data_ptr(), buf["ws_bytes"], stream.cuda_stream))

    def _capture(self, L, i, out_dev):
        torch = self.torch
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = L.ocrpp_launch_count()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._launch(L, torch.cuda.current_stream(), out_dev)
            self._graphs[i] = (g, int(L.ocrpp_launch_count() - n0))
            return self._graphs[i]
        except Exception as e:   # capture not possible here: direct launches
            sys.stderr.write("[bench] CUDA graph capture of the DB step failed (%s): direct launches\n" % (e,))
            self._graph_failed = True
            try:
                torch.cuda.synchronize()
            except Exception:
                pass
            return None


class DbFp16Workload(DbWorkload):
    name = "db_fp16"
    dtype = "f16"
    map_dtype = "float16"
    elem_bytes = 2
    stream_kernel = "db_scan4_kernel<__half>"
    workload = ("DB++ r18 post-process on fp16 probability maps (SURVEY 8(d) variant), batch 256 synthetic 736x1280 maps "
                "per GPU")


class ExpandWorkload(DetWorkload):
    def e2e_step(self):
        # pinned HOST maps straight into the operator (chunked upload overlapped with the kernels)
        return self.op({"maps": self.host_maps}, self.shape_list)

    def extra_cfg(self):
        return {"maps_at_processing_res": True}

    def device_step(self, L, stream):
        from pytorchocr_b200 import _lib
        op, buf, m = self.op, self.buf, self.dev_maps
        _, N, C, h, w, fin, cap, R, arena = self.key
        i, out_dev, _ = self.out_pair(stream)
        _lib.check(op._call_lib(L, m, N, C, h, w, fin, op.scale, dict(buf, out_dev=out_dev), cap, R, arena, None, None, stream))
        self.out_copy(i, stream)


class PseWorkload(ExpandWorkload):
    name = "pse"
    default_batch = 16
    channels = 7
    op_name = "PSEPostProcess"
    cfg = PSE_CFG
    stream_kernel = "ex_binarize_kernel"
    workload = ("PSENet r50 progressive scale expansion, 7 kernel maps at processing resolution 736x1280, "
                "~200 text regions each, batch 16 per GPU = 128 over 8 B200 (BASELINE.json configs[2])")

    def host_prepare(self, pool, want_cpu):
        self.scenes = pool.map(_gen_pse_scene, [SEED + self.rank * self.batch + i for i in range(self.batch)])
        if not want_cpu:
            return None
        n = max(self.cores, 8)
        out, dt = _timed_pool(pool, _cpu_pse, [SEED + i for i in range(n)], self.cores)
        return {"value": n / dt, "unit": self.unit, "cores": self.cores, "kind": "port",
                "sample": "%d maps [7,736,1280] from the same generator in %.1f s, multiprocessing.Pool(%d); "
                          "oracle/pse_oracle.py (C restatement of pse.pyx pinned against the compiled reference + "
                          "numpy/cv2 generate_box without the reference's O(labels*H*W) scans); mean %.2f s/image/core"
                          % (n, dt, self.cores, float(np.mean([o[0] for o in out])))}

    def make_device_maps(self, dev):
        from pytorchocr_b200 import synth
        return synth.pse_maps_torch(self.scenes, SEED + self.rank, dev)


class PanWorkload(ExpandWorkload):
    name = "pan"
    default_batch = 128
    channels = 6
    op_name = "PANPostProcess"
    cfg = PAN_CFG
    stream_kernel = "ex_binarize_kernel"
    workload = ("PAN++ r18 pixel aggregation with 4-d embeddings, [6,736,1280] maps at processing resolution, "
                "~200 text regions each, batch 128 per GPU (BASELINE.json configs[3])")

    def host_prepare(self, pool, want_cpu):
        self.scenes = pool.map(_gen_pan_scene, [SEED + self.rank * self.batch + i for i in range(self.batch)])
        if not want_cpu:
            return None
        n = max(self.cores, 8)
        out, dt = _timed_pool(pool, _cpu_pan, [SEED + i for i in range(n)], self.cores)
        return {"value": n / dt, "unit": self.unit, "cores": self.cores, "kind": "port",
                "sample": "%d maps [6,736,1280] from the same generator in %.1f s, multiprocessing.Pool(%d); "
                          "oracle/pan_oracle.py (numpy pre-pass + C restatement of pa.pyx pinned against the compiled "
                          "reference + generate_box without the O(labels*H*W) scans); mean %.2f s/image/core"
                          % (n, dt, self.cores, float(np.mean([o[0] for o in out])))}

    def make_device_maps(self, dev):
        from pytorchocr_b200 import synth
        return synth.pan_maps_torch(self.scenes, SEED + self.rank, dev)

    def alg_bytes_per_unit(self):
        return 6 * self.in_h * self.in_w * 4

    def kernel_bytes_per_unit(self):
        # only the text and kernel channels are read unconditionally (the 4 embedding channels are
        # touched for flagged kernels only), so the streaming kernel is rated on the 2 channels it must read while
        # whole_step_frac keeps the 6-channel rule
        return 2 * self.in_h * self.in_w * 4


def _gen_pse_scene_q(seed):
    from pytorchocr_b200 import synth
    return synth.pse_scene(seed, H // 4, W // 4, n_abs=200, hh_rng=(2, 3), hw_rng=(4, 9))[:2]


def _gen_pan_scene_q(seed):
    from pytorchocr_b200 import synth
    return synth.pan_scene(seed, H // 4, W // 4, n_abs=200, hh_rng=(2, 3), hw_rng=(4, 9))[:4]


class PseApiWorkload(PseWorkload):
    """SURVEY 8(d) API-faithful twin: the head's own output [N,7,184,320], scale 1 -> the operator's nearest x4
    up-sampling is fused into the first kernel; processing resolution 736x1280."""
    name = "pse_api"
    in_h, in_w = H // 4, W // 4
    cfg = dict(PSE_CFG, scale=1)
    workload = ("PSENet r50 post-process through the operator API: head output [16,7,184,320] per GPU, scale 1 "
                "(x4 nearest up-sampling fused, processing resolution 736x1280)")

    def extra_cfg(self):
        return {}

    def host_prepare(self, pool, want_cpu):
        self.scenes = pool.map(_gen_pse_scene_q, [SEED + self.rank * self.batch + i for i in range(self.batch)])
        return None


class PanApiWorkload(PanWorkload):
    """SURVEY 8(d) API-faithful twin: head output [N,6,184,320], scale 4 (the shipped det_r18_pan.yml): processing at
    184x320, labels up-sampled x4 for the boxes."""
    name = "pan_api"
    in_h, in_w = H // 4, W // 4
    cfg = dict(PAN_CFG, scale=4)
    workload = ("PAN++ r18 post-process through the operator API: head output [128,6,184,320] per GPU, scale 4 "
                "(processing resolution 184x320, boxes at 736x1280)")

    def extra_cfg(self):
        return {}

    def host_prepare(self, pool, want_cpu):
        self.scenes = pool.map(_gen_pan_scene_q, [SEED + self.rank * self.batch + i for i in range(self.batch)])
        return None

    def alg_bytes_per_unit(self):
        return 6 * self.in_h * self.in_w * 4

    def kernel_bytes_per_unit(self):
        return 2 * self.in_h * self.in_w * 4


class CtcWorkload(Workload):
    name = "ctc"
    unit = "lines/s"
    metric = "CTC greedy decode text lines/sec (T=80, 6623 classes)"
    default_batch = 8192
    stream_kernel = "ctc_argmax_kernel"
    workload = ("CRNN vgg CTC greedy decode, T=80, 6623 classes, 8192 text lines (17.4 GB) per GPU = 65536 over "
                "8 B200 (BASELINE.json configs[4])")
    e2e_lines = 512

    def alg_bytes_per_unit(self):
        return CTC_T * CTC_C * 4

    def host_prepare(self, pool, want_cpu):
        import tempfile
        from pytorchocr_b200 import synth
        self.dict_path = synth.write_char_dict(os.path.join(tempfile.mkdtemp(), "dict.txt"), CTC_C)
        if not want_cpu:
            return None
        per, n = 128, self.cores * 2
        out, dt = _timed_pool(pool, _cpu_ctc, [(SEED + i, per, self.dict_path) for i in range(n)], self.cores)
        inner = sum(o[0] for o in out)
        return {"value": per * n / (inner / self.cores), "unit": self.unit, "cores": self.cores, "kind": "port",
                "sample": "%d chunks of %d lines [80,%d,6623] from the same generator, multiprocessing.Pool(%d), decode "
                          "time only (%.1f s summed); oracle/ctc_oracle.py CTCLabelDecodeNumpy = the reference's numpy "
                          "argmax + pure-Python collapse loop restated line by line" % (n, per, per, self.cores, inner)}

    def device_prepare(self, dev):
        import torch
        from pytorchocr_b200 import synth
        from pytorchocr_b200.postprocess import build_post_process
        self.torch, self.dev = torch, dev
        self.op = build_post_process({"name": "CTCLabelDecode", "character_dict_path": self.dict_path,
                                      "use_space_char": False, "cuda_speedup": True})
        B = self.batch
        self.probs = synth.ctc_probs_torch(SEED + self.rank, CTC_T, B, CTC_C, dev)
        self.idx = torch.empty(B * CTC_T + B, dtype=torch.int32, device=dev)
        self.pf = torch.empty(B * CTC_T + B, dtype=torch.float32, device=dev)
        self.idx_host = torch.empty(B * CTC_T + B, dtype=torch.int32, pin_memory=True)
        self.conf_host = torch.empty(B, dtype=torch.float32, pin_memory=True)
        e = min(self.e2e_lines, B)
        self.host_chunk = torch.empty((CTC_T, e, CTC_C), dtype=torch.float32, pin_memory=True)
        self.host_chunk.copy_(self.probs[:, :e])
        self.n_boxes = 0
        self.op(self.probs[:, :e])

    def device_step(self, L, stream):
        from pytorchocr_b200 import _lib
        B, x = self.batch, self.probs
        _lib.check(L.ocrpp_ctc_greedy(x.data_ptr(), _lib.F32, CTC_T, B, CTC_C, x.stride(0), x.stride(1),
                                      self.idx.data_ptr(), self.pf.data_ptr(), self.idx.data_ptr() + 4 * B * CTC_T,
                                      self.pf.data_ptr() + 4 * B * CTC_T, None, stream.cuda_stream))
        self.idx_host.copy_(self.idx, non_blocking=True)
        self.conf_host.copy_(self.pf[B * CTC_T:], non_blocking=True)

    def e2e_step(self):
        d = self.host_chunk.to(self.dev, non_blocking=True)
        return self.op(d)     # kernels + D2H of ids/lengths/confidences + host string assembly

    e2e_units = property(lambda self: self.host_chunk.shape[1])

    def h2d_bytes(self):
        return int(self.host_chunk.numel() * 4)

    def d2h_bytes(self):
        e = self.host_chunk.shape[1]
        return int((e * CTC_T + e) * 4 + e * 4)

    def public_config(self):
        return {"batch_per_gpu": self.batch, "T": CTC_T, "C": CTC_C,
                "l2": "inputs (%.1f GB per GPU) larger than the 126 MB L2; no flush needed"
                      % (self.batch * self.alg_bytes_per_unit() / 1e9)}

    def detail(self):
        return {"timed_region": "device probabilities -> kept class ids / lengths / confidences in pinned host memory",
                "e2e_region": "%d-line chunks: pinned host probabilities -> python strings" % self.host_chunk.shape[1]}


class CropWorkload(Workload):
    """the text-line crops between the detector and the recogniser."""
    name = "crop"
    unit = "images/s"
    metric = "text-line crop pages/sec @736x1280 (200 boxes per page)"
    default_batch = 64
    dtype = "u8"
    stream_kernel = "crop_warp_kernel"
    stream_phase = 1
    workload = ("sort_boxes + get_part_img + rot90 rule for 200 detected boxes on each of 64 synthetic 736x1280x3 uint8 "
                "pages per GPU (the step after BASELINE.json configs[1]'s detector path)")

    def alg_bytes_per_unit(self):
        # every crop pixel written once and the source pixels under it read once
        return 2 * self.crop_bytes // self.batch

    def host_prepare(self, pool, want_cpu):
        seeds = [SEED + self.rank * 1000 + i for i in range(self.batch)]
        self.pages = pool.map(_gen_page, seeds, chunksize=2)
        if not want_cpu:
            return None
        sample = self.pages[:max(16, self.cores)] * 4
        out, dt = _timed_pool(pool, _cpu_crop, sample, self.cores)
        inner = sum(o[0] for o in out)
        return {"value": len(sample) / (inner / self.cores), "unit": self.unit, "cores": self.cores, "kind": "port",
                "sample": "%d pages of this step (4 passes), multiprocessing.Pool(%d), cv2.setNumThreads(1), crop time "
                          "only (%.2f s summed); oracle/crop_oracle.py = the reference's own sequence of cv2 calls "
                          "(utility.py:32-78)" % (len(sample), self.cores, inner)}

    def device_prepare(self, dev):
        import torch
        from pytorchocr_b200.part_img import PartImageCropper
        self.torch, self.dev = torch, dev
        self.op = PartImageCropper()
        self.host_imgs = torch.from_numpy(np.stack([p[0] for p in self.pages])).pin_memory()
        self.host_boxes = torch.from_numpy(np.stack([p[1] for p in self.pages])).pin_memory()
        self.imgs = self.host_imgs.to(dev)
        self.boxes = self.host_boxes.to(dev)
        self.counts = torch.full((self.batch,), CROP_BOXES, dtype=torch.int32, device=dev)
        arena, offsets, dims, order, status = self.op.run_device(self.imgs, self.boxes, self.counts)
        assert not status.any(), status
        self.crop_bytes = int(offsets[-1])
        self.n_boxes = self.batch * CROP_BOXES
        self.buf = next(iter(self.op._bufs.values()))
        K = self.batch * CROP_BOXES
        self.meta_host = torch.empty(self.buf["meta"].numel(), dtype=torch.int32, pin_memory=True)
        self.crops_host = torch.empty(self.crop_bytes, dtype=torch.uint8, pin_memory=True)
        self.K = K

    def device_step(self, L, stream):
        from pytorchocr_b200 import _lib
        b, K = self.buf, self.K
        base = b["meta"].data_ptr()
        o_dims = base + 8 * (K + 1)
        _lib.check(L.ocrpp_crop_boxes(self.imgs.data_ptr(), self.batch, H, W, 3, self.imgs.stride(0), self.imgs.stride(1),
                                      self.boxes.data_ptr(), self.counts.data_ptr(), CROP_BOXES, 1, 1,
                                      b["arena"].data_ptr(), b["arena"].numel(), base, o_dims, o_dims + 8 * K,
                                      o_dims + 12 * K, b["ws"].data_ptr(), b["ws"].numel(), stream.cuda_stream))
        self.meta_host.copy_(b["meta"], non_blocking=True)      # offsets / dims / order / status; the crops stay in HBM

    def e2e_step(self):
        imgs = self.host_imgs.to(self.dev, non_blocking=True)
        boxes = self.host_boxes.to(self.dev, non_blocking=True)
        arena, offsets, dims, order, status = self.op.run_device(imgs, boxes, self.counts)
        self.crops_host.copy_(arena[:self.crop_bytes], non_blocking=True)     # the crops themselves come back
        self.torch.cuda.current_stream().synchronize()
        return offsets

    def h2d_bytes(self):
        return int(self.host_imgs.numel() + self.host_boxes.numel() * 2)

    def d2h_bytes(self):
        return int(self.crop_bytes + self.meta_host.numel() * 4)

    def public_config(self):
        return {"batch_per_gpu": self.batch, "H": H, "W": W, "boxes_per_page": CROP_BOXES,
                "l2": "pages (%.0f MB per GPU) larger than the 126 MB L2; no flush needed" % (self.batch * H * W * 3 / 1e6)}

    def detail(self):
        return {"crop_bytes_per_step": self.crop_bytes,
                "timed_region": "device pages + device boxes -> crops in a device arena (where the recogniser reads "
                                "them), offsets/dims/order in pinned host memory",
                "e2e_region": "pinned host pages + boxes -> crops in pinned host memory"}


WORKLOADS = {"db": DbWorkload, "pse": PseWorkload, "pan": PanWorkload, "ctc": CtcWorkload, "crop": CropWorkload,
             "db_fp16": DbFp16Workload, "pse_api": PseApiWorkload, "pan_api": PanApiWorkload}
OTHERS = ["pse", "pan", "ctc", "pse_api", "pan_api", "db_fp16"]     # reported under other_workloads by the default run


# ----------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path on the host cores
# ----------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    wl = WORKLOADS[args.workload](args, 0, 1)
    note = ("the oracle port of the reference's CPU path (oracle/): db_postprocess.cpp needs OpenCV C++ and cannot be "
            "built here, pse.pyx/pa.pyx are restated in C and pinned against the compiled reference")
    if args.workload in ("db", "db_fp16"):
        # the SAME maps our arm processes on rank 0 (same generator, same seeds, same batch)
        with mp.get_context("fork").Pool(cores) as pool:
            maps = np.stack(pool.map(_gen_db, [SEED + i for i in range(wl.batch)], chunksize=4))[:, None]
        for _ in range(max(0, args.warmup - 1)):
            _cpu_db_batch(maps, cores)
        rate, nboxes, dt = _cpu_db_batch(maps, cores, passes=args.steps)
        units, what, note = wl.batch, "all %d maps of the step (%d boxes per pass)" % (wl.batch, nboxes // max(1, args.steps)), DbWorkload.cpu_note
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            if args.workload in ("pse", "pan", "pse_api", "pan_api"):
                sample = max(8, cores)
                items = [SEED + i for i in range(sample)]
                fn, units = (_cpu_pse if args.workload.startswith("pse") else _cpu_pan), sample
                what = "%d maps at processing resolution" % sample
            elif args.workload == "crop":
                sample = max(16, cores)
                items = pool.map(_gen_page, [SEED + i for i in range(sample)], chunksize=2) * 8
                fn, units = _cpu_crop, sample * 8
                what = "%d pages" % (sample * 8)
                note = "oracle/crop_oracle.py: the reference's own sequence of cv2 calls (utility.py:32-78)"
            else:
                import tempfile
                from pytorchocr_b200 import synth
                dict_path = synth.write_char_dict(os.path.join(tempfile.mkdtemp(), "dict.txt"), CTC_C)
                per, n = 128, cores * 2
                items = [(SEED + i, per, dict_path) for i in range(n)]
                fn, units = _cpu_ctc, per * n
                what = "%d chunks of %d lines" % (n, per)
            pool.map(fn, items[:cores])
            for _ in range(args.warmup):
                pool.map(fn, items, chunksize=1)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                pool.map(fn, items, chunksize=1)
            dt = time.perf_counter() - t0
        rate = units * args.steps / dt
    line = {
        "impl": "reference", "metric": wl.metric, "value": rate, "unit": wl.unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
        "config": dict({"workload": wl.workload}, **wl.public_config()),
        "cpu_baseline": {"value": rate, "unit": wl.unit, "cores": cores, "kind": "port",
                         "sample": "%s x %d steps, fork pool of %d workers, cv2.setNumThreads(1); %s"
                                   % (what, args.steps, cores, note)},
        "e2e": {"value": rate, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def _pin_rank_to_cores(local_rank, world):
    """Each rank gets its own slice of the host cores (the ranks of one node otherwise migrate over each other's
    cores while they enqueue kernels, which shows up as per-rank jitter of the event-timed window)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(1, world))
        mine = cores[local_rank * per:(local_rank + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return len(mine)
    except Exception:
        return None


def _traffic(wl):
    """roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum of the streaming kernel from one
    `ncu --set full` capture (profiles/traffic.json holds it PER UNIT, with the capture it came from), scaled to the
    units of one launch of this run."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[wl.name]
        return float(t["dram_bytes_per_unit"]) * wl.batch
    except Exception:
        return None


def measure(wl, args, steps, L, torch, dist, dev, local_rank, world, sampler):
    """Device-resident timing, per-kernel phases and end-to-end timing of one prepared workload -> dict."""
    from pytorchocr_b200 import _lib
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if wl.needs_flush() else None
    for _ in range(max(3, args.warmup)):
        wl.device_step(L, stream)
    barrier()
    launches0 = L.ocrpp_launch_count() + getattr(wl, "replayed_launches", 0)
    barrier()
    t_enq = None
    if flush is None:
        # inputs larger than the L2: K steps back to back between one pair of events
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.start()
        e0.record(stream)
        t_h0 = time.perf_counter()
        for _ in range(steps):
            wl.device_step(L, stream)
        if hasattr(wl, "device_join"):
            wl.device_join(stream)
        e1.record(stream)
        t_enq = (time.perf_counter() - t_h0) * 1e3 / steps     # host time to ENQUEUE one step (not to run it)
        barrier()
        sampler.stop()
        ms_dev = e0.elapsed_time(e1)
    else:
        # small inputs: the L2 is flushed before every step; each step has its own pair of events
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        sampler.start()
        for a, b in evs:
            flush.fill_(1)
            a.record(stream)
            wl.device_step(L, stream)
            if hasattr(wl, "device_join"):
                wl.device_join(stream)
            b.record(stream)
        barrier()
        sampler.stop()
        ms_dev = sum(a.elapsed_time(b) for a, b in evs)
    launches = L.ocrpp_launch_count() + getattr(wl, "replayed_launches", 0) - launches0
    # ---- per-kernel durations (events around every kernel inside the library), in a separate pass: with the marks
    #      on, the library runs each batch as ONE chain on the caller's stream, so the phases add up to an
    #      un-overlapped step; the timed region above is the production configuration ----
    wl.graph_ok = False      # the event marks live in the library's host code: direct launches for this pass
    L.ocrpp_profile_enable(1)
    for _ in range(2):
        wl.device_step(L, stream)
    barrier()
    L.ocrpp_profile_reset()
    for _ in range(min(steps, 5)):
        if flush is not None:
            flush.fill_(1)
        wl.device_step(L, stream)
    barrier()
    L.ocrpp_profile_enable(0)
    wl.graph_ok = True
    calls, phases = _lib.profile_read()
    # ---- end-to-end timing through the operator with HOST buffers ----
    for _ in range(2):
        wl.e2e_step()
    barrier()
    e2e_steps = max(1, min(args.e2e_steps, 20) if wl.name == "db" else min(steps, 5))
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        wl.e2e_step()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    sampler.stop()

    t = torch.tensor([ms_dev, t_e2e * 1e3, t_enq if t_enq is not None else -1.0], dtype=torch.float64, device=dev)
    mine = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev_max, ms_e2e_max = float(t[0]), float(t[1])
    rank_ms, rank_e2e = [ms_dev / steps], [wl_units(wl) * e2e_steps / t_e2e]
    rank_enq = [t_enq]
    if world > 1:   # per-rank step times next to the max the headline uses
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_ms = [float(x[0]) / steps for x in allr]
        rank_e2e = [wl_units(wl) * e2e_steps / (float(x[1]) * 1e-3) for x in allr]
        rank_enq = [float(x[2]) if float(x[2]) >= 0 else None for x in allr]

    # per-kernel durations of every rank (the kernels are the same; a rank whose GPU runs one of them slower shows here)
    phases_by_rank = None
    if world > 1 and phases:
        pv = torch.tensor([ms / max(1, calls) for _, ms in phases], dtype=torch.float64, device=dev)
        allp = [torch.zeros_like(pv) for _ in range(world)]
        dist.all_gather(allp, pv)
        phases_by_rank = {nm: [round(float(x[i]), 4) for x in allp] for i, (nm, _) in enumerate(phases)}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = ("MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks
                else "fallback 6.65 TB/s (B200_PROFILING.md)")
    alg_bytes = wl.batch * wl.alg_bytes_per_unit()
    kernel_bytes = wl.batch * (wl.kernel_bytes_per_unit() if hasattr(wl, "kernel_bytes_per_unit") else wl.alg_bytes_per_unit())
    sp = getattr(wl, "stream_phase", 0)
    k1_ms = phases[sp][1] / max(1, calls) if len(phases) > sp else None
    achieved = kernel_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms else None
    cfg = {"workload": wl.workload}
    cfg.update(wl.public_config())
    return {
        "metric": wl.metric, "value": world * wl.batch * steps / (ms_dev_max * 1e-3), "unit": wl.unit,
        "steps": steps, "ms_per_step": ms_dev_max / steps, "dtype": wl.dtype, "config": cfg, "detail": wl.detail(),
        "roofline": {"bound": "hbm", "kernel": wl.stream_kernel, "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": _traffic(wl),
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                     "kernel_bytes_per_launch": kernel_bytes, "kernel_ms": k1_ms,
                     "whole_step_frac": alg_bytes / (ms_dev_max / steps * 1e-3) / 1e9 / peak},
        "phases_ms": {nm: ms / max(1, calls) for nm, ms in phases},
        "e2e": {"value": world * wl_units(wl) * e2e_steps / (ms_e2e_max * 1e-3), "unit": wl.unit,
                "h2d_bytes_per_step": wl.h2d_bytes(), "d2h_bytes_per_step": wl.d2h_bytes(), "steps": e2e_steps,
                "by_rank": rank_e2e},
        "gpu_launches": int(launches),
        "ms_per_step_by_rank": rank_ms,
        "host_enqueue_ms_per_step_by_rank": rank_enq,
        "phases_ms_by_rank": phases_by_rank,
        "cuda_graph": bool(getattr(wl, "_graphs", None)) and not getattr(wl, "_graph_failed", False),
    }


def wl_units(wl):
    return getattr(wl, "e2e_units", wl.batch)


def h2d_probe(torch, dist, dev, local_rank, world):
    """Plain pinned host->device copy rate of every rank, all ranks copying at the same time: the ceiling the
    end-to-end numbers are read against (8 ranks share one host's memory controllers and PCIe root complexes)."""
    n = 256 << 20
    src = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    dst = torch.empty(n, dtype=torch.uint8, device=dev)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[local_rank])
    t0 = time.perf_counter()
    for _ in range(4):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    gbs = torch.tensor([4 * n / (time.perf_counter() - t0) / 1e9], dtype=torch.float64, device=dev)
    out = [gbs.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(out, gbs)
    return [float(x[0]) for x in out]


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    pinned = _pin_rank_to_cores(local_rank, world) if world > 1 else None
    names = [args.workload] + ([n for n in OTHERS if n != args.workload]
                               if (args.workload == "db" and not args.headline_only and not args.batch) else [])
    wls = [WORKLOADS[n](args, rank, world) for n in names]

    # ---- host phase (fork pool; CUDA not initialised yet) ----
    import multiprocessing as mp
    cpu_base = None
    for i, wl in enumerate(wls):
        with mp.get_context("fork").Pool(wl.cores) as pool:
            cb = wl.host_prepare(pool, i == 0 and rank == 0 and world == 1 and not args.no_cpu)
        if i == 0:
            cpu_base = cb

    # ---- device phase ----
    import torch
    import torch.distributed as dist
    from pytorchocr_b200 import _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    sampler = ClockSampler(local_rank)
    results = []
    for i, wl in enumerate(wls):
        wl.device_prepare(dev)
        steps = args.steps if i == 0 else max(3, min(args.steps, 5))
        results.append(measure(wl, args, steps, L, torch, dist, dev, local_rank, world, sampler))
        wl.release()
    h2d = h2d_probe(torch, dist, dev, local_rank, world)
    # every rank's own clock summary (a rank that is slower than the others at identical kernels is usually a GPU
    # that sat at a lower SM clock under its power cap): median SM MHz and a bit mask of the reasons seen
    mine = sampler.summary()
    bits = sum(b for nm, b in (("hw_slowdown", 1), ("hw_thermal_slowdown", 2), ("sw_thermal_slowdown", 4), ("sw_power_cap", 8))
               if nm in mine.get("reasons", []))
    ck = torch.tensor([mine.get("sm_mhz") or 0.0, float(bits)], dtype=torch.float64, device=dev)
    cks = [ck.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(cks, ck)

    if rank == 0:
        head = results[0]
        line = {
            "metric": head["metric"], "value": head["value"], "unit": head["unit"],
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": head["dtype"], "data": "synthetic", "config": head["config"],
            "detail": head["detail"],
            "roofline": head["roofline"], "phases_ms": head["phases_ms"], "e2e": head["e2e"],
            "gpu_launches": head["gpu_launches"], "ms_per_step_by_rank": head["ms_per_step_by_rank"],
            "host_enqueue_ms_per_step_by_rank": head["host_enqueue_ms_per_step_by_rank"], "cuda_graph": head["cuda_graph"],
            "phases_ms_by_rank": head["phases_ms_by_rank"],
            "h2d_gbs_by_rank": h2d, "host_cores_per_rank": pinned,
            "clocks": sampler.summary(),
            "sm_mhz_by_rank": [float(x[0]) for x in cks],
            "clock_reason_bits_by_rank": [int(x[1]) for x in cks],   # 1 hw_slowdown 2 hw_thermal 4 sw_thermal 8 sw_power_cap
        }
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        if len(results) > 1:
            line["other_workloads"] = {
                wl.name: {k: r[k] for k in ("metric", "value", "unit", "steps", "ms_per_step", "dtype", "config", "detail", "roofline",
                                            "phases_ms", "e2e", "gpu_launches", "ms_per_step_by_rank",
                                            "host_enqueue_ms_per_step_by_rank", "cuda_graph", "phases_ms_by_rank")}
                for wl, r in zip(wls[1:], results[1:])}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="db", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="units per GPU per step (0 = the workload's default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--headline-only", action="store_true", help="skip the other_workloads of the default run")
    ap.add_argument("--e2e-steps", type=int, default=20, help="end-to-end steps of the headline workload")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and world == 1:
        # not launched by torchrun: re-exec one rank per GPU on this node
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000),
               os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
