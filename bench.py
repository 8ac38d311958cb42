#!/usr/bin/env python3
"""Headline benchmark: Mvoxels/s of the igneous hot path (downsample 2 mode mips
-> 6-connected CCL -> marching-cubes meshing at mip 2) on a synthetic 2048^3
uint32 segmentation resident in HBM, one z-slab of the dataset per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl reference]

Contract (see the task brief): W untimed warm-up steps, exactly K timed steps
bracketed by barrier + device synchronisation, CUDA-event timing on the stream
the kernels are launched on, max over ranks, ONE JSON line from rank 0.
"""
import argparse
import ctypes as c
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

def _baseline_metric():
  """The metric string of BASELINE.json (the bench line must name exactly that metric)."""
  fallback = "Mvoxels/s on 2048\u00b3 uint32 seg (downsample+CCL+mesh) @1/2/4/8 B200; % HBM roofline"
  try:
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "BASELINE.json")) as f:
      return json.load(f).get("metric", fallback)
  except (OSError, ValueError):
    return fallback


METRIC = _baseline_metric()
RESOLUTION = (16, 16, 40)
PITCH, NUM_IDS = 64, 1 << 20


def measured_peaks():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    try:
      return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
      pass
  return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
  """nvidia-smi clocks / throttle reasons during the timed region."""

  def __init__(self, index):
    super().__init__(daemon=True)
    self.index, self.samples, self.stop_flag = index, [], False
    self.proc = None

  def run(self):
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
      self.proc = subprocess.Popen(
        ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
         "-lms", "500"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      for line in self.proc.stdout:
        if self.stop_flag:
          break
        self.samples.append([v.strip() for v in line.split(",")])
    except Exception:
      pass

  def finish(self):
    self.stop_flag = True
    if self.proc is not None:
      try:
        self.proc.kill()
      except Exception:
        pass
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for s in self.samples:
      try:
        sm.append(float(s[0]))
        mx.append(float(s[1]))
        for n, v in zip(names, s[3:7]):
          if v.lower().startswith("active"):
            reasons.add(n)
      except Exception:
        continue
    return {"sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------- CPU legs
def oracle_pipeline(seg, simplify=100):
  """The CPU restatement of one step on a host array: returns voxels processed."""
  from oracle import oracle as O
  mips = O.downsample_segmentation(seg, (2, 2, 1), num_mips=2)
  O.connected_components(seg, out_dtype=np.uint32)
  m2 = mips[1]
  for z0 in range(0, m2.shape[2], 256):
    for y0 in range(0, m2.shape[1], 256):
      for x0 in range(0, m2.shape[0], 256):
        tl, tv = O.marching_cubes(m2[x0:x0 + 257, y0:y0 + 257, z0:z0 + 257])
        W = O.WeldedMeshes(tl, tv)
        if simplify:
          O.simplify_welded(W, RESOLUTION, simplify, 40.0, True)
  return seg.size


_WORKER = {}
REF_CHUNK = (256, 256, 64)


def host_cores():
  """Cores this process may actually use: the affinity mask clipped by the cgroup CPU quota
  (os.cpu_count() reports the whole node even inside a small lease)."""
  try:
    n = len(os.sched_getaffinity(0))
  except (AttributeError, OSError):
    n = os.cpu_count() or 1
  quota = None
  try:
    with open("/sys/fs/cgroup/cpu.max") as f:  # cgroup v2
      q, per = f.read().split()[:2]
      if q != "max":
        quota = float(q) / float(per)
  except (OSError, ValueError):
    try:
      with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f1, open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f2:
        q, per = float(f1.read()), float(f2.read())
        if q > 0:
          quota = q / per
    except (OSError, ValueError):
      pass
  if quota is not None:
    n = max(1, min(n, int(quota + 0.5)))
  return n


def ref_chunk_offset(index, size):
  """Chunk `index` of the REF_CHUNK grid over the size^3 bench volume (x fastest, as
  FinelyDividedTaskIterator enumerates tasks, igneous/task_creation/common.py:91-98)."""
  gx, gy, gz = (max(1, size // c) for c in REF_CHUNK)
  index %= gx * gy * gz
  return ((index % gx) * REF_CHUNK[0], ((index // gx) % gy) * REF_CHUNK[1], (index // (gx * gy)) * REF_CHUNK[2])


def _oracle_worker_init(counter, size):
  """Each pool worker synthesises its own chunk of the bench volume ONCE (outside any timed region)."""
  from oracle import oracle as O
  with counter.get_lock():
    wid = counter.value
    counter.value += 1
  # spread the workers' chunks over the volume (stride 37 is coprime with the chunk grid)
  _WORKER["seg"] = O.synth_seg(REF_CHUNK, pitch=PITCH, num_ids=NUM_IDS, seed=0,
                               offset=ref_chunk_offset(wid * 37, size))
  O.lib()


def _oracle_worker(_):
  t = time.perf_counter()
  n = oracle_pipeline(_WORKER["seg"])
  return n, time.perf_counter() - t


def run_reference_arm(args):
  """--impl reference: the reference's CPU implementation of the path.  The
  reference's own kernels (tinybrain / cc3d / zmesh wheels) are absent from
  this image, so this times the C oracle port on the host cores this process may
  use (one chunk worker per core, spawn, as igneous_cli/cli.py:915-933 does).  Every
  worker holds one 256x256x64 chunk cut from the SAME synthetic volume the GPU arm
  processes; a step = every worker runs the pipeline once on its chunk."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  import multiprocessing as mp
  from oracle import oracle as O
  O.build()
  cores = host_cores()
  ctx = mp.get_context("spawn")
  # one chunk alone on an otherwise idle host: the reference's real per-worker speed
  _oracle_worker_init(ctx.Value("i", 0), args.size)
  _oracle_worker(0)
  alone = min(_oracle_worker(0)[1] for _ in range(2))
  counter = ctx.Value("i", 0)
  times, per_chunk = [], []
  with ctx.Pool(cores, initializer=_oracle_worker_init, initargs=(counter, args.size)) as pool:
    pool.map(_oracle_worker, range(cores), chunksize=1)  # untimed: all workers initialised and warm
    for it in range(args.warmup + args.steps):
      t = time.perf_counter()
      res = pool.map(_oracle_worker, range(cores), chunksize=1)
      dt = time.perf_counter() - t
      if it >= args.warmup:
        times.append((sum(r[0] for r in res), dt))
        per_chunk.extend(r[1] for r in res)
  vox = sum(t[0] for t in times)
  sec = sum(t[1] for t in times)
  value = vox / sec / 1e6
  shape = REF_CHUNK
  line = {
    "impl": "reference", "metric": METRIC, "value": value, "unit": "Mvoxels/s", "n_gpus": args.gpus,
    "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / max(len(times), 1),
    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
    "data": "synthetic", "gpu_launches": 0,
    "config": {"workload": "oracle port of the igneous CPU path on %dx%dx%d uint32 chunks cut from the %d^3 "
                           "jittered-Voronoi bench volume (pitch 64), one chunk per usable host core (%d) per step: "
                           "mode pool 2 mips + 6-connected CCL + marching cubes / weld / quadric simplification "
                           "x100 at mip 2" % (shape + (args.size, cores)),
               "chunk": list(shape), "simplification_factor": 100, "cores_used": cores,
               "os_cpu_count": os.cpu_count(),
               "seconds_per_chunk_alone": alone,
               "seconds_per_chunk_contended_median": float(np.median(per_chunk)) if per_chunk else None},
    "cpu_baseline": {"value": value, "unit": "Mvoxels/s", "cores": cores, "kind": "port",
                     "sample": "%d x %dx%dx%d chunks per step, %d steps" % ((cores,) + shape + (args.steps,))},
    "e2e": {"value": value, "unit": "Mvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }
  print(json.dumps(line))


def cpu_baseline_sample(pipe, ctx, budget_s=20.0):
  """Oracle timed on ONE host core on a bounded sample of the same volume."""
  from oracle import oracle as O
  O.build()
  sx, sy, sz = pipe.shape
  bz = min(sz, 256)
  bx, by = min(sx, 256), min(sy, 256)
  from igneous_b200 import _shim
  d_box = ctx.alloc(bx * by * bz * 4)
  _shim.check(ctx.lib.ign_copy_box_dev(ctx.handle, _shim.ptr(pipe.d_in), c.c_int(pipe.code),
                                       c.c_uint64(sx), c.c_uint64(sy), c.c_uint64(sz), c.c_uint64(0),
                                       c.c_uint64(0), c.c_uint64(0), c.c_uint64(bx), c.c_uint64(by),
                                       c.c_uint64(bz), _shim.ptr(d_box)))
  seg = ctx.to_host(d_box, (bx, by, bz), np.uint32)
  d_box.free()
  t = time.perf_counter()
  vox, reps = 0, 0
  while True:
    vox += oracle_pipeline(seg, pipe.simplification_factor)
    reps += 1
    if time.perf_counter() - t > budget_s / 2 or reps >= 8:
      break
  dt = time.perf_counter() - t
  return {"value": vox / dt / 1e6, "unit": "Mvoxels/s", "cores": 1, "kind": "port",
          "sample": "%dx%dx%d corner of the bench volume, %d repetitions, %.1f s" % (bx, by, bz, reps, dt)}


# -------------------------------------------------------------------- main
def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=5)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--size", type=int, default=2048, help="cube edge of the per-GPU volume")
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--no-e2e", action="store_true")
  ap.add_argument("--no-cpu", action="store_true")
  ap.add_argument("--e2e-steps", type=int, default=2)
  ap.add_argument("--e2e-shared-buffers", action="store_true",
                  help="download the labels into the input host buffer (forced automatically when host RAM is tight)")
  ap.add_argument("--simplify", type=int, default=None, help="simplification factor (default 100)")
  ap.add_argument("--mesh-streams", type=int, default=8, help="concurrent MeshTask bodies per GPU")
  ap.add_argument("--serial-simplify", action="store_true",
                  help="use the serial ring walkers of the simplifier instead of the batched-gather kernels")
  ap.add_argument("--ccl-v1", action="store_true",
                  help="use k_ccl_local_fast instead of the 4-voxels-per-lane tile kernel k_ccl_local_v2")
  args = ap.parse_args()
  if args.warmup < 3 and args.impl == "b200":
    args.warmup = 3

  if args.impl == "reference":
    return run_reference_arm(args)

  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  dist = None
  if world > 1:
    import torch
    import torch.distributed as dist_mod
    torch.cuda.set_device(local_rank)
    dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dist = dist_mod

  # The simplifier's batched-gather kernels (simplify.cu, IGN_SIMP_BATCH) give bit-identical
  # meshes (tests/test_mesh_gpu.py::test_simplify_batched_kernels_bit_exact) and 28 % shorter
  # MeshTask bodies; the library default is still the serial kernels until the whole GPU suite
  # has run with them, so the bench opts in explicitly.
  if args.serial_simplify:
    os.environ.pop("IGN_SIMP_BATCH", None)
  else:
    os.environ.setdefault("IGN_SIMP_BATCH", "1")
  # Same for the CCL tile kernel: k_ccl_local_v2 (ccl.cu, IGN_CCL_V2) is bit-exact
  # (tests/test_ccl_gpu.py::test_ccl_v2_kernel_matches_oracle) and 11 % faster at 512^3.
  if args.ccl_v1:
    os.environ.pop("IGN_CCL_V2", None)
  else:
    os.environ.setdefault("IGN_CCL_V2", "1")
  from igneous_b200 import _shim, pipeline
  ctx = _shim.Context(local_rank)
  S = args.size
  shape = (S, S, S)
  simplify = 100 if args.simplify is None else args.simplify
  if not hasattr(ctx.lib, "ign_mesh_simplify"):
    simplify = 0
  group = None
  if world > 1:
    from igneous_b200 import multigpu
    group = multigpu.Group(ctx, rank, world, dist)
  pipe = pipeline.VolumePipeline(ctx, shape, np.uint32, num_mips=2, mesh_shape=(256, 256, 256),
                                 resolution=RESOLUTION, pitch=PITCH, num_ids=NUM_IDS, seed=0,
                                 offset=(0, 0, rank * S), simplification_factor=simplify, group=group,
                                 mesh_streams=args.mesh_streams)
  pipe.synth()
  ctx.sync()

  def barrier():
    ctx.sync()
    if dist is not None:
      dist.barrier()

  for _ in range(args.warmup):
    pipe.step(timers=False)
  barrier()

  sampler = ClockSampler(local_rank) if rank == 0 else None
  if sampler:
    sampler.start()
    time.sleep(0.3)
  launches0 = pipe.launch_count()
  pipe.prof_enable(True)
  stage = {"pool_ms": 0.0, "ccl_ms": 0.0, "mesh_ms": 0.0}
  barrier()
  ctx.timer_start(0)
  for _ in range(args.steps):
    pipe.step(timers=True)
    for k, v in pipe.stage_ms().items():
      stage[k] += v
  ctx.timer_stop(0)
  total_ms = ctx.timer_ms(0)
  barrier()
  prof = pipe.prof_read()
  pipe.prof_enable(False)
  launches = pipe.launch_count() - launches0
  clocks = sampler.finish() if sampler else None

  if dist is not None:
    import torch
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
  ms_per_step = total_ms / args.steps
  voxels = pipe.n * world
  value = voxels / (ms_per_step * 1e-3) / 1e6

  # ---- roofline of the dominant kernel group (CCL), SURVEY.md 8(d):
  # algorithmic bytes = in_bytes + out_bytes per voxel of the CCL stage
  peak, peak_src = measured_peaks()
  in_b, out_b = 4, pipe.ccl_out_dtype.itemsize
  ccl_ms = sum(prof[k][0] for k in ("ccl_local", "ccl_merge", "ccl_label")) / args.steps
  kern = {k: {"ms_per_step": prof[k][0] / args.steps, "launches_per_step": prof[k][1] / args.steps}
          for k in prof}
  dominant = max(("ccl_local", "ccl_merge", "ccl_label", "pool"), key=lambda k: prof[k][0])
  alg = {"ccl_local": in_b + 4, "ccl_merge": 0, "ccl_label": 4 + out_b, "pool": 4 * (1 + 0.25 + 0.0625)}
  dom_ms = prof[dominant][0] / args.steps
  dom_launches = max(prof[dominant][1] / args.steps, 1)
  dom_bytes_per_launch = pipe.n * alg[dominant] / dom_launches
  achieved = (dom_bytes_per_launch / 1e9) / (dom_ms / dom_launches * 1e-3) if dom_ms > 0 else 0.0
  roofline = {
    "bound": "hbm", "kernel": "k_" + dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
    "frac": achieved / peak,
    # dram__bytes_read+write of this kernel from the ncu --set full capture in
    # profiles/r01_ccl_local_fast_full_512_raw.csv: 7.657 B/voxel (537+491 MB at 512^3), scaled per launch
    "traffic": (7.657 * pipe.n / dom_launches) if dominant == "ccl_local" else None,
    "traffic_source": "ncu capture at 512^3 scaled by voxels per launch (profiles/%s; both tile kernels move 7.65 B/voxel)"
                      % ("r01_ccl_local_v2_full_512_raw.csv" if os.environ.get("IGN_CCL_V2") else "r01_ccl_local_fast_full_512_raw.csv"),
    "peak_source": peak_src,
    "algorithmic_bytes_per_voxel": alg[dominant],
    "avg_launch_ms": dom_ms / dom_launches, "launches_per_step": dom_launches,
    "stage_ccl": {"algorithmic_bytes_per_voxel": in_b + out_b, "ms_per_step": ccl_ms,
                  "achieved": pipe.n * (in_b + out_b) / 1e9 / (ccl_ms * 1e-3) if ccl_ms > 0 else 0.0},
    "kernels": kern,
    "note": "dominant own streaming kernel; the mesh stage (cub sorts + latency-bound simplification rounds on "
            "%d concurrent streams) takes %.0f%% of the step and has no bandwidth roofline"
            % (pipe.mesh_streams, 100.0 * stage["mesh_ms"] / max(sum(stage.values()), 1e-9)),
  }
  roofline["stage_ccl"]["frac"] = roofline["stage_ccl"]["achieved"] / peak

  line = {
    "metric": METRIC, "value": value, "unit": "Mvoxels/s", "n_gpus": world, "steps": args.steps,
    "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
    "vs_baseline": None, "dtype": "u32", "data": "synthetic", "gpu_launches": int(launches),
    "config": {
      "workload": "%d^3 uint32 jittered-Voronoi segmentation per GPU (pitch %d): mode-pool 2 mips, "
                  "6-connected CCL at mip 0 (u32 ids), marching cubes + weld%s at mip 2 in 256^3 tasks"
                  % (S, PITCH, (" + quadric simplification x%d" % simplify) if simplify else
                     " (simplification NOT in the timed region: kernel not landed yet)"),
      "volume_per_gpu": list(shape), "parallelism": "z-slab per GPU, %d rank(s)" % world,
      "l2": "inputs larger than L2 (%.1f GB volume vs 126 MB L2)" % (pipe.n * 4 / 1e9),
      "simplification_factor": simplify, "components": pipe.n_components, "mesh": pipe.mesh_stats, "mesh_streams": pipe.mesh_streams,
      "simplify_kernels": "batched gathers (IGN_SIMP_BATCH=1)" if os.environ.get("IGN_SIMP_BATCH") else "serial ring walks",
      "ccl_tile_kernel": "k_ccl_local_v2 (IGN_CCL_V2=1)" if os.environ.get("IGN_CCL_V2") else "k_ccl_local_fast",
      "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()},
    },
    "roofline": roofline, "clocks": clocks,
  }

  if rank == 0 and not args.no_cpu:
    line["cpu_baseline"] = cpu_baseline_sample(pipe, ctx)
  if not args.no_e2e:
    line["e2e"] = run_e2e(ctx, pipe, args, dist, world)
  if rank == 0:
    print(json.dumps(line))
  if dist is not None:
    dist.barrier()
    dist.destroy_process_group()


def _mem_available():
  try:
    with open("/proc/meminfo") as f:
      for line in f:
        if line.startswith("MemAvailable:"):
          return int(line.split()[1]) * 1024
  except OSError:
    pass
  return None


def run_e2e(ctx, pipe, args, dist, world):
  """Same metric through host buffers: every step copies the volume H2D from
  pinned memory, runs the pipeline and copies every product (mips, CCL labels,
  all mesh fragments) back D2H."""
  from igneous_b200 import _shim
  n = pipe.n
  host_kind = "pinned (cudaHostAlloc)"
  # host RAM guard: all local ranks keep input + products resident on the host.  When that
  # does not fit comfortably, the label volume is downloaded into the input buffer (what
  # in_place=True does in the reference) and the input is restored between steps, untimed.
  out_bytes = sum(int(np.prod(s)) * 4 for s in pipe.mip_shapes) + n * pipe.ccl_out_dtype.itemsize
  local_ranks = int(os.environ.get("LOCAL_WORLD_SIZE", world))
  avail = _mem_available()
  if dist is not None:  # every rank must take the same decisions below: agree on the smallest reading
    import torch
    t = torch.tensor([float(avail) if avail is not None else -1.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    avail = int(t.item()) if t.item() >= 0 else None
  shared = (avail is not None and (n * 4 + out_bytes) * local_ranks > 0.6 * avail
            and pipe.ccl_out_dtype.itemsize == 4)
  if args.e2e_shared_buffers:
    shared = pipe.ccl_out_dtype.itemsize == 4
  need = (n * 4 + (out_bytes - n * pipe.ccl_out_dtype.itemsize if shared else out_bytes)) * local_ranks
  if avail is not None and need > 0.85 * avail:
    # never drive the host out of memory: report the leg as not measurable on this box
    return {"value": None, "unit": "Mvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
            "skipped": "host buffers for %d local rank(s) need %.0f GB, %.0f GB of host RAM available"
                       % (local_ranks, need / 1e9, avail / 1e9)}
  try:
    host_in = ctx.pinned_empty(pipe.shape, np.uint32)
    host = {"mips": [ctx.pinned_empty(s, np.uint32) for s in pipe.mip_shapes],
            "cc": (host_in.view(pipe.ccl_out_dtype) if shared
                   else ctx.pinned_empty(pipe.shape, pipe.ccl_out_dtype))}
  except (MemoryError, _shim.IgneousB200Error):
    # the host could not page-lock the buffers for this rank: fall back to pageable memory
    host_kind = "pageable (pinned allocation failed)"
    host_in = np.empty(pipe.shape, dtype=np.uint32, order="F")
    host = {"mips": [np.empty(s, dtype=np.uint32, order="F") for s in pipe.mip_shapes],
            "cc": (host_in.view(pipe.ccl_out_dtype) if shared
                   else np.empty(pipe.shape, dtype=pipe.ccl_out_dtype, order="F"))}
  ctx.d2h(host_in, pipe.d_in)
  cap_v, cap_f = 1 << 22, 1 << 23
  ctx.sync()
  mesh_bytes = [0]
  res = (c.c_float * 3)(*[float(r) for r in RESOLUTION])

  lock = threading.Lock()
  host_bufs = {}

  def export(task, h, nv, nf, nl, wctx):
    if nv == 0:
      return
    if id(wctx) not in host_bufs:  # one pinned staging pair per mesh stream
      host_bufs[id(wctx)] = (wctx.pinned_empty((cap_v, 3), np.float32, order="C"),
                             wctx.pinned_empty((cap_f, 3), np.uint32, order="C"))
    bv, bf = host_bufs[id(wctx)]
    voff = np.zeros(nl + 1, dtype=np.uint64)
    foff = np.zeros(nl + 1, dtype=np.uint64)
    v = bv if nv <= cap_v else np.empty((nv, 3), np.float32)
    f = bf if nf <= cap_f else np.empty((nf, 3), np.uint32)
    _shim.check(wctx.lib.ign_mesh_export(h, res, c.c_int(1), _shim.ptr(v), _shim.ptr(f),
                                         _shim.ptr(voff), _shim.ptr(foff)))
    with lock:
      mesh_bytes[0] += nv * 12 + nf * 12

  steps = max(1, min(args.e2e_steps, args.steps))

  def one():
    mesh_bytes[0] = 0
    # upload in z-layers; pooling, meshing, CCL and the D2H of the products overlap
    pipe.step_streamed(host_in, host, export)

  one()  # warm-up (pinned pages touched, arena sized)
  ctx.sync()
  ms = 0.0
  for _ in range(steps):
    if shared:  # the label download overwrote the input: restore it outside the timed region
      ctx.d2h(host_in, pipe.d_in)
      ctx.sync()
    if dist is not None:
      dist.barrier()
    t0 = time.perf_counter()
    ctx.timer_start(4)
    one()
    ctx.timer_stop(4)
    ev = ctx.timer_ms(4)
    ms += max(ev, (time.perf_counter() - t0) * 1e3)  # host-side export work counts too
  if dist is not None:
    import torch
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
  h2d = n * 4
  d2h = sum(int(np.prod(s)) * 4 for s in pipe.mip_shapes) + n * pipe.ccl_out_dtype.itemsize + mesh_bytes[0]
  return {"value": n * world / (ms / steps * 1e-3) / 1e6, "unit": "Mvoxels/s",
          "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": steps,
          "ms_per_step": ms / steps, "host_memory": host_kind,
          "host_buffers": ("input buffer reused for the label download, restored between steps outside "
                           "the timed region" if shared else "separate input and output buffers")}


if __name__ == "__main__":
  main()
