"""Benchmark of the RigL hot path: sparse ResNet-50 train step + mask update.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): sparse train-step images/sec, ResNet-50, 80 % ERK, bf16,
batch 256 per GPU (configs[1]; weak scaling: 256 images per GPU at every N),
synthetic ImageNet-shaped data, RigL schedule drop 0.3 / cosine / every 100
steps, so a 100-step timed region contains exactly one mask update.
One JSON line on rank 0; keys per the driver contract plus `roofline`,
`cpu_baseline`, `mask_update_ms`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 256
IMAGE = 224
SPARSITY = 0.8
METRIC = 'sparse_train_step_images_per_sec_resnet50_erk80'
# masked FLOPs per image, 2*MAC, maskable layers only.  This build computes the
# DENSE wgrad every step (as the TF reference effectively does), so the "update step"
# accounting 2*f_S + f_D - f_S(first conv) applies to every step.
ALG_GFLOP_PER_IMAGE = 14.744
DENSE_GFLOP_PER_IMAGE = 3 * 8.178 - 0.236       # dense-executed fprop+dgrad+wgrad, no stem dgrad


def _peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as f:
      d = json.load(f)
    return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops_sustained', 1400.0), 'measured'
  return 6650.0, 1400.0, 'fallback'


def _recorded_traffic():
  """DRAM bytes per step of the conv kernel family from the committed ncu pass
  (profiles/r01_dram_traffic_step.json: dram__bytes_read.sum + dram__bytes_write.sum, b256)."""
  path = os.path.join(ROOT, 'profiles', 'r01_dram_traffic_step.json')
  try:
    with open(path) as f:
      d = json.load(f)
    return {'dram_bytes_per_step': d['conv_family_dram_bytes_per_step'], 'launches': d['conv_family_launches'],
            'source': 'profiles/r01_dram_traffic_step.json'}
  except Exception:
    return None


class ClockSampler(object):
  """Samples nvidia-smi clocks / throttle reasons during the timed region."""

  def __init__(self, gpu_index=0):
    self.rows, self.proc, self.thread, self.idx = [], None, None, gpu_index

  def start(self):
    q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + q,
                                    '--format=csv,noheader,nounits', '-lms', '200'],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
      return
    self.thread = threading.Thread(target=self._pump, daemon=True)
    self.thread.start()

  def _pump(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(',')])

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except Exception:
      self.proc.kill()
    sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
    mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == 'Active' for r in self.rows)]
    return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
            'reasons': reasons, 'samples': len(sm)}


def _dist_setup(n_gpus):
  import torch.distributed as dist
  world = int(os.environ.get('WORLD_SIZE', '1'))
  if world > 1:
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    return dist, dist.get_rank(), world, local
  torch.cuda.set_device(0)
  return None, 0, 1, 0


def run_ours(args):
  from rigl_b200 import _cabi
  from rigl_b200 import workloads
  from rigl_b200.layers import Profiler

  dist, rank, world, local = _dist_setup(args.gpus)
  dev = torch.device('cuda', local)
  torch.manual_seed(0)
  model = workloads.ResNet50(device=dev)
  workloads.init_masks(model, 'erdos_renyi_kernel', SPARSITY, seed=0)
  dp = None
  if world > 1:
    from rigl_b200.data_parallel import DataParallel
    dp = DataParallel()
  harness = workloads.TrainHarness(model, lr=0.1, data_parallel=dp)
  g = torch.Generator(device=dev).manual_seed(1 + rank)
  images = torch.randn(BATCH, 3, IMAGE, IMAGE, device=dev, generator=g).to(torch.bfloat16) \
      .contiguous(memory_format=torch.channels_last)
  labels = torch.randint(0, 1000, (BATCH,), device=dev, generator=g)

  def barrier():
    if dist is not None:
      dist.barrier()
    torch.cuda.synchronize()

  harness.step(images, labels)               # eager: first step is the initial mask update
  harness.step(images, labels)
  graphed = False
  if not args.no_graph:
    graphed = harness.enable_cuda_graph(images, labels)
  for _ in range(args.warmup):
    harness.step(images, labels)
  barrier()
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  launches0 = _cabi.launch_count() + getattr(harness, 'replayed_kernel_launches', 0)
  start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  start.record()
  n_updates = 0
  for _ in range(args.steps):
    harness.step(images, labels)
    n_updates += int(harness.opt.last_update_was_mask_update)
  stop.record()
  barrier()
  clocks = sampler.stop() if rank == 0 else None
  launches = _cabi.launch_count() + getattr(harness, 'replayed_kernel_launches', 0) - launches0
  ms = torch.tensor([start.elapsed_time(stop)], device=dev, dtype=torch.float64)
  if dist is not None:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  total_ms = float(ms.item())
  value = world * BATCH * args.steps / (total_ms / 1e3)

  # ---- end-to-end leg: host (pinned) -> device copy of every batch, loss read back ----
  e2e_steps = max(3, min(args.steps, 20))
  host_images = torch.empty((BATCH, IMAGE, IMAGE, 3), dtype=torch.bfloat16).pin_memory()
  host_images.copy_(images.permute(0, 2, 3, 1).cpu())
  host_labels = labels.cpu().pin_memory()
  barrier()
  e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  copy_stream = torch.cuda.Stream(device=dev)

  def fetch():            # host -> device copy of one batch on the copy stream (input prefetch)
    with torch.cuda.stream(copy_stream):
      xb = host_images.to(dev, non_blocking=True)
      yb = host_labels.to(dev, non_blocking=True)
      ev = torch.cuda.Event()
      ev.record(copy_stream)
    return xb, yb, ev

  e_start.record()
  nxt = fetch()
  for i in range(e2e_steps):
    xb, yb, ev = nxt
    torch.cuda.current_stream().wait_event(ev)
    if i + 1 < e2e_steps:
      nxt = fetch()       # overlaps the next batch's H2D with this step's compute
    loss = harness.step(xb.permute(0, 3, 1, 2), yb)
    xb.record_stream(torch.cuda.current_stream())
    _ = float(loss.item())
  e_stop.record()
  barrier()
  e_ms = torch.tensor([e_start.elapsed_time(e_stop)], device=dev, dtype=torch.float64)
  if dist is not None:
    dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
  e2e_value = world * BATCH * e2e_steps / (float(e_ms.item()) / 1e3)

  # ---- roofline leg: per-call CUDA-event times of the conv kernels (all ranks step: the
  # data-parallel all-reduce is collective; only rank 0 records) ----
  prof_steps = 3
  harness.graphed = False                    # the per-call event timing needs the eager path
  for _ in range(2):                         # re-warm the eager allocator state after graph replay
    harness.step(images, labels)
  barrier()
  if rank == 0:
    Profiler.start()
  for _ in range(prof_steps):
    harness.step(images, labels)
  barrier()
  if rank != 0:
    if dist is not None:
      dist.destroy_process_group()
    return
  rec = Profiler.stop()
  per_kind = {}
  for kind, _, t in rec:
    per_kind[kind] = per_kind.get(kind, 0.0) + t / prof_steps
  if args.layer_report:
    agg = {}
    for kind, scope, t in rec:
      agg[(kind, scope)] = agg.get((kind, scope), 0.0) + t / prof_steps
    with open(args.layer_report, 'w') as f:
      json.dump([{'kind': k, 'scope': sc, 'ms': v} for (k, sc), v in agg.items()], f, indent=0)
  conv_ms = sum(per_kind.get(k, 0.0) for k in ('fprop', 'dgrad', 'wgrad'))
  n_conv_launch = sum(1 for k, _, _ in rec if k in ('fprop', 'dgrad', 'wgrad')) / prof_steps
  hbm_peak, tf_peak, peak_src = _peaks()
  achieved_tf = ALG_GFLOP_PER_IMAGE * BATCH / conv_ms            # GFLOP/ms == TFLOP/s
  # ---- mask update alone (all 54 layers, one update) ----
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
  mu = []
  harness.opt.drop_fraction = np.float32(0.3)
  for i in range(8):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    harness.opt.mask_update_op()
    e.record()
    torch.cuda.synchronize()
    if i >= 3:
      mu.append(s.elapsed_time(e))
  mask_ms = float(np.median(mu))
  total_w = sum(m.size for m in model.registry.get_masks())

  out = {
      'metric': METRIC, 'value': value, 'unit': 'images/sec', 'n_gpus': world, 'steps': args.steps,
      'warmup': args.warmup, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
      'config': {'workload': 'ResNet-50 ImageNet-shaped, 80% ERK (54 masked tensors, 25.5M weights), '
                             'batch 256/GPU, RigL drop 0.3 cosine every 100 steps, Nesterov momentum',
                 'global_batch': BATCH * world, 'parallelism': 'dp%d' % world,
                 'l2_policy': 'inputs larger than L2 (activations per step >> 126 MB)',
                 'mask_updates_in_timed_region': n_updates,
                 'cuda_graph': bool(graphed)},
      'clocks': clocks,
      'e2e': {'value': e2e_value, 'unit': 'images/sec', 'steps': e2e_steps,
              'h2d_bytes_per_step': int(host_images.numel() * 2 + host_labels.numel() * 8),
              'd2h_bytes_per_step': 4},
      'gpu_launches': int(launches),
      'mask_update_ms': mask_ms,
      'mask_update_algorithmic_GBps': 8.25 * total_w / mask_ms / 1e6,
      'roofline': {'bound': 'tensor', 'kernel': 'k_igemm_kmajor2 / k_igemm_wgrad / k_halo3x3_* (all masked conv+linear launches)',
                   'achieved': achieved_tf, 'peak': tf_peak, 'unit': 'TFLOP/s', 'frac': achieved_tf / tf_peak,
                   'peak_source': peak_src + ' bf16_tflops_sustained',
                   'algorithmic_gflop_per_image': ALG_GFLOP_PER_IMAGE,
                   'dense_executed_tflops': DENSE_GFLOP_PER_IMAGE * BATCH / conv_ms,
                   'conv_ms_per_step': conv_ms, 'conv_launches_per_step': n_conv_launch,
                   'ms_per_step_by_kind': per_kind, 'traffic': _recorded_traffic()},
  }
  if world == 1 and not args.no_cpu_baseline:
    out['cpu_baseline'] = cpu_baseline_leg(sample_batch=args.cpu_batch)
  _emit(out)
  if dist is not None:
    dist.destroy_process_group()


def _cpu_port_timing(batch, steps, warmup):
  """Times the CPU port in a FRESH interpreter whose OpenMP environment is not the one
  torchrun exports (OMP_NUM_THREADS=1): torch then sizes its intra-op pool to the host's
  cores.  Returns {'sec_per_step', 'mask_update_sec', 'threads'}."""
  env = dict(os.environ)
  for k in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OMP_PROC_BIND', 'OMP_PLACES', 'GOMP_CPU_AFFINITY',
            'KMP_AFFINITY', 'CUDA_VISIBLE_DEVICES'):
    env.pop(k, None)
  env['CUDA_VISIBLE_DEVICES'] = ''
  code = ('import json,sys,torch; sys.path.insert(0, %r); '
          'from oracle import cpu_train_step as c; '
          's, net, dense = c.time_train_steps(%d, %d, warmup=%d); '
          'mu = c.time_mask_update(net, dense); '
          'print("CPUPORT " + json.dumps({"sec_per_step": s, "mask_update_sec": mu, '
          '"threads": torch.get_num_threads()}))' % (ROOT, batch, steps, warmup))
  out = subprocess.run([sys.executable, '-c', code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=1500)
  for line in out.stdout.splitlines():
    if line.startswith('CPUPORT '):
      return json.loads(line[len('CPUPORT '):])
  raise RuntimeError('CPU port failed: ' + out.stderr[-2000:])


def cpu_baseline_leg(sample_batch=16, steps=1):
  """Times the CPU port of the reference path on the host cores (bounded sample)."""
  t = _cpu_port_timing(sample_batch, steps, 1)
  return {'value': sample_batch / t['sec_per_step'], 'unit': 'images/sec', 'cores': t['threads'], 'kind': 'port',
          'sample': 'ResNet-50 80%% ERK fp32 train step (fwd + dense&masked bwd + momentum), batch %d, '
                    '%d timed step(s) after 1 warm-up, torch-CPU port of the TF1 graph' % (sample_batch, steps),
          'mask_update_ms': t['mask_update_sec'] * 1e3,
          'mask_update_sample': 'one drop/grow update of all 54 layers (numpy stable argsort x2 per layer)'}


def run_reference(args):
  """The reference's own CPU implementation of the path (torch-CPU / numpy port of the
  TF1 graph -- TensorFlow is not installable in this image), all host threads."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  batch = args.cpu_batch
  steps = max(1, min(args.steps, 3))
  warm = max(1, min(args.warmup, 1))
  t0 = time.perf_counter()
  t = _cpu_port_timing(batch, steps, warm)
  sec, mu, threads = t['sec_per_step'], t['mask_update_sec'], t['threads']
  value = batch / sec
  _emit({
      'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'images/sec',
      'n_gpus': args.gpus, 'steps': steps, 'warmup': warm, 'ms_per_step': sec * 1e3,
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': 'ResNet-50 ImageNet-shaped, 80%% ERK, CPU port of the reference TF1 train step, '
                             'bounded sample of batch %d per step' % batch},
      'cpu_baseline': {'value': value, 'unit': 'images/sec', 'cores': threads, 'kind': 'port',
                       'sample': 'batch %d, %d step(s), wall %.1fs' % (batch, steps, time.perf_counter() - t0)},
      'mask_update_ms': mu * 1e3,
      'e2e': {'value': value, 'unit': 'images/sec', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'gpu_launches': 0})


_JSON_FD = None


def _emit(obj):
  """The ONE JSON line of the contract, on the process's original stdout."""
  line = (json.dumps(obj) + '\n').encode()
  if _JSON_FD is None:
    sys.stdout.write(line.decode())
    sys.stdout.flush()
  else:
    os.write(_JSON_FD, line)


def main():
  # stdout carries exactly one JSON line: everything else that writes to fd 1 (NCCL prints its
  # version banner there when NCCL_DEBUG is set, library warnings) is sent to stderr.
  global _JSON_FD
  sys.stdout.flush()
  _JSON_FD = os.dup(1)
  os.dup2(2, 1)
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=100)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--cpu-batch', type=int, default=16)
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--layer-report', default=None)
  ap.add_argument('--no-graph', action='store_true', help='run the step eagerly (no CUDA-graph replay)')
  args = ap.parse_args()
  if args.warmup < 3 and args.impl == 'ours':
    args.warmup = 3
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
