#!/usr/bin/env python
"""Benchmark of the dGPMP2 inner Gauss-Newton step (BASELINE.json metric, config[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path (PlanLayer.forward: factors -> block-tridiagonal normal
equations -> solve -> dtheta, err, err_ext) over one batch of B = 1024 synthetic 2-D point-robot
problems with T = 64 states and a 128 x 128 SDF each ("2D point robot, batch=1024 random-obstacle
envs, 64 states, 1xB200").  With N GPUs every rank processes its own B problems (the planning
batch is sharded; no data-path collective), so the job is weak-scaled and
value = N * B * K / max-over-ranks(time).

Printed JSON (one line, rank 0):
  value      device-resident throughput (inputs already in HBM), CUDA-event timed, K launches
  e2e        same metric through the host-buffer C-ABI entry (dgpmp2_gn_step_host_f32): pinned host
             inputs -> H2D -> kernel -> D2H of dtheta/err/status, every step, copies inside the timed region
  roofline   algorithmic bytes per launch / average launch duration, against the measured HBM peak
  cpu_baseline  the CPU oracle port (oracle/gn_oracle.py = the reference's dense algorithm) timed on
             this box's host cores on a bounded sample of the same workload
--impl reference times that CPU port alone (rank 0 only) and prints the same line shape.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU = 1024
T_STATES = 64
IM_SIZE = 128
N_SETS = 4                 # rotating input sets: 4 x (64 MiB SDF + 1 MiB th) = 260 MiB > 126 MB L2
ITERATE = 5                # the timed iterate: 5 GN updates after the straight-line initialisation
YAML = dict(Q_c_inv=[[1.0, 0.0], [0.0, 1.0]], K_s=0.01, K_g=0.01, cost_sigma=0.01, epsilon_dist=0.4,
            reg=0.1, total_time_sec=10.0, sphere_radius=0.4)
METRIC = 'gn_problem_iters_per_sec'
UNIT = 'problem-iters/s'


def algorithmic_bytes(B, T, d=4, es=4):
    """th in + dth out, start + goal, 4 SDF taps per state, err + err_ext (static weights)."""
    return B * (2 * T * d * es + 2 * d * es + 16 * T + 8)


def make_cparams():
    from dgpmp2_b200 import _lib
    return _lib.make_params(B=B_PER_GPU, T=T_STATES, dof=2, H=IM_SIZE, W=IM_SIZE, x_lims=(-5.0, 5.0), y_lims=(-5.0, 5.0),
                            total_time_sec=YAML['total_time_sec'], r_sphere=YAML['sphere_radius'], K_s=YAML['K_s'],
                            K_g=YAML['K_g'], reg=YAML['reg'], Q_c_inv=YAML['Q_c_inv'], cost_sigma=YAML['cost_sigma'],
                            epsilon_dist=YAML['epsilon_dist'])


def make_inputs(seed, n_sets, B):
    from dgpmp2_b200.datasets.synthetic import make_problems
    sets = []
    for s in range(n_sets):
        pr = make_problems(B, T_STATES, dof=2, im_size=IM_SIZE, seed=1000 * seed + s, unique_envs=256, dtype=torch.float32)
        sets.append(pr)
    return sets


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the benchmark runs (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) >= 8:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        busy = [r for r in self.rows if r[3].isdigit() and int(r[3]) > 0] or self.rows
        for r in busy:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for k, n in enumerate(names):
                if r[4 + k].lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'samples': len(busy),
                'reasons': sorted(reasons)}


def cpu_oracle_setup(sample_B, pr):
    from oracle import gn_oracle
    p = gn_oracle.GNParams(dof=2, T=T_STATES, total_time_sec=YAML['total_time_sec'], x_lims=[-5.0, 5.0], y_lims=[-5.0, 5.0],
                           r_sphere=YAML['sphere_radius'], K_s=YAML['K_s'], K_g=YAML['K_g'], reg=YAML['reg'],
                           Q_c_inv=YAML['Q_c_inv'], cost_sigma=YAML['cost_sigma'], epsilon_dist=YAML['epsilon_dist'])
    th = pr['th_init'][:sample_B].double()
    start, goal, sdf = pr['start'][:sample_B].double(), pr['goal'][:sample_B].double(), pr['sdf'][:sample_B].double()
    qc = torch.tensor(YAML['Q_c_inv'], dtype=torch.float64).expand(sample_B, T_STATES - 1, 2, 2)
    w = torch.full((sample_B, T_STATES, 1, 1), 1.0 / YAML['cost_sigma'] ** 2, dtype=torch.float64)
    eps = torch.full((sample_B, T_STATES, 1, 1), YAML['epsilon_dist'], dtype=torch.float64)

    def step():
        return gn_oracle.gn_step(th, start, goal, sdf, qc, w, eps, p)
    return step


def time_cpu(step, steps, warmup, sample_B):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return sample_B * steps / dt, dt / steps


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_worker(args):
    """Child process: time the CPU oracle port and print one JSON line (kept in a child with a hard
    timeout so that a misbehaving host BLAS can never hang the benchmark)."""
    torch.set_num_threads(host_threads())
    sample_B = args.cpu_sample
    pr = make_inputs(0, 1, sample_B)[0]
    step = cpu_oracle_setup(sample_B, pr)
    val, sec = time_cpu(step, args.steps, args.warmup, sample_B)
    print(json.dumps({'value': val, 'sec_per_step': sec, 'threads': torch.get_num_threads(), 'sample_B': sample_B,
                      'steps': args.steps, 'warmup': args.warmup}))


def run_cpu_worker(steps, warmup, sample_B, timeout_s):
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'cpu-worker', '--steps', str(steps), '--warmup', str(warmup),
           '--cpu-sample', str(sample_B)]
    env = dict(os.environ)
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'MASTER_ADDR', 'MASTER_PORT'):
        env.pop(k, None)
    try:
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=timeout_s, env=env)
        for line in reversed(res.stdout.strip().splitlines()):
            if line.startswith('{'):
                return json.loads(line), None
        return None, 'cpu worker produced no result (rc=%d)' % res.returncode
    except subprocess.TimeoutExpired:
        return None, 'cpu worker exceeded %d s' % timeout_s


def cpu_baseline_obj(r):
    return {'value': r['value'], 'unit': UNIT, 'cores': r['threads'], 'kind': 'port',
            'sample': '%d problems x %d steps of the same synthetic workload (dense A/b/K + dense Cholesky + explicit '
                      'triangular inverses in torch fp64, %.3f s/step)' % (r['sample_B'], r['steps'], r['sec_per_step'])}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference is pure Python and cannot
    travel to the GPU box) on this box's host cores, all threads, bounded sample per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample_B = 128
    steps = max(1, min(args.steps, 10))
    warm = max(1, min(args.warmup, 2))
    r, why = run_cpu_worker(steps, warm, sample_B, 240)
    if r is None:
        print(json.dumps({'impl': 'reference', 'unavailable': why}))
        return
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': warm, 'ms_per_step': r['sec_per_step'] * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': '2D point robot, batch=1024 random-obstacle envs, 64 states (CPU: bounded sample of %d problems per step)' % sample_B,
                   'batch_per_step': sample_B, 'states': T_STATES, 'sdf': '%dx%d' % (IM_SIZE, IM_SIZE)},
        'cpu_baseline': cpu_baseline_obj(r),
        'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20000)
    ap.add_argument('--warmup', type=int, default=50)
    ap.add_argument('--impl', type=str, default='b200')
    ap.add_argument('--e2e-steps', type=int, default=20)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-sample', type=int, default=128)
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.impl == 'cpu-worker':
        return cpu_worker(args)

    from dgpmp2_b200 import _lib, ops
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU baseline')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    W = max(3, args.warmup)
    K = max(1, args.steps)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    B, T, d = B_PER_GPU, T_STATES, 4
    lib = _lib.load()
    cp = make_cparams()
    sets = make_inputs(rank, N_SETS, B)
    dsets = []
    for pr in sets:
        th, start, goal, sdf = (pr[k].to(dev).contiguous() for k in ('th_init', 'start', 'goal', 'sdf'))
        # advance to the timed iterate with the persistent solver (same kernels, untimed)
        th_it = ops.gn_solve(cp, th, start, goal, sdf, ITERATE, 0.0)[0]
        dsets.append((th_it.contiguous(), start.reshape(B, d).contiguous(), goal.reshape(B, d).contiguous(), sdf[:, 0].contiguous()))
    dth = torch.empty(B, T, d, device=dev)
    err = torch.empty(B, device=dev)
    err_ext = torch.empty(B, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    _lib.set_sdf_shape(cp, IM_SIZE, IM_SIZE, IM_SIZE * IM_SIZE)
    cp.B = B
    fn = lib.dgpmp2_gn_step_f32
    pref = ctypes.byref(cp)
    vp = ctypes.c_void_p
    argsets = [(vp(a.data_ptr()), vp(b.data_ptr()), vp(c.data_ptr()), vp(s.data_ptr())) for (a, b, c, s) in dsets]
    outs = (vp(dth.data_ptr()), vp(err.data_ptr()), vp(err_ext.data_ptr()), vp(status.data_ptr()))
    stream = torch.cuda.current_stream()
    sp = vp(stream.cuda_stream)

    def launch(i):
        a = argsets[i % N_SETS]
        rc = fn(pref, a[0], a[1], a[2], a[3], None, outs[0], outs[1], outs[2], outs[3], sp)
        if rc != 0:
            _lib.check(rc)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ----------------
    # The K launches are captured in CUDA graphs (chunks of <= 100 launches over the rotating input
    # sets) and replayed, so the timed region contains exactly K kernel launches back to back and no
    # Python / ctypes work between them.
    for i in range(W):
        launch(i)
    barrier()
    chunk = min(K, 100)
    n_full, rem = divmod(K, chunk)

    def capture(n, first):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            gs = vp(torch.cuda.current_stream().cuda_stream)
            for i in range(n):
                a = argsets[(first + i) % N_SETS]
                rc = fn(pref, a[0], a[1], a[2], a[3], None, outs[0], outs[1], outs[2], outs[3], gs)
                if rc != 0:
                    _lib.check(rc)
        return g
    g_full = capture(chunk, 0)
    g_rem = capture(rem, 0) if rem else None
    g_full.replay()                      # warm the instantiated graph
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n_full):
        g_full.replay()
    if g_rem is not None:
        g_rem.replay()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    assert int(status.abs().max()) == 0, 'factorisation failure in the timed region'
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * B * K / (ms_max * 1e-3)

    # ---------------- end to end with host buffers ----------------
    Ke = max(1, min(args.e2e_steps, K))
    hs = ops.HostStepper(cp, torch.float32, dev)
    hsets = []
    for pr in sets[:2]:
        hsets.append(tuple(pr[k].contiguous().pin_memory() for k in ('th_init', 'start', 'goal', 'sdf')))

    def e2e_run(resident):
        for i in range(3):
            h = hsets[i % len(hsets)]
            hs.step(h[0], h[1], h[2], h[3], sdf_resident=False)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a0.record(stream)
        for i in range(Ke):
            h = hsets[0] if resident else hsets[i % len(hsets)]
            out = hs.step(h[0], h[1], h[2], h[3], sdf_resident=resident)
            _ = float(out[1][0])                       # the step's result (err) is read on the host
        a1.record(stream)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([max(a0.elapsed_time(a1), wall)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * B * Ke / (float(t.item()) * 1e-3)

    e2e_val = e2e_run(False)
    e2e_res = e2e_run(True)

    # ---------------- SURVEY 8(d) extras (device-resident, not the headline): iterate 0 / 10, persistent solve ----------------
    def time_launches(fn_one, n):
        for _ in range(3):
            fn_one()
        barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n):
                fn_one()
        g.replay()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        g.replay()
        a1.record(stream)
        barrier()
        return a0.elapsed_time(a1) / n

    extras = {}
    pr0 = sets[0]
    th0, st0, go0, sdf0 = (pr0[k].to(dev).contiguous() for k in ('th_init', 'start', 'goal', 'sdf'))
    for k_it in (0, 10):
        thk = th0 if k_it == 0 else ops.gn_solve(cp, th0, st0, go0, sdf0, k_it, 0.0)[0].contiguous()
        a = (vp(thk.data_ptr()), vp(st0.reshape(B, d).data_ptr()), vp(go0.reshape(B, d).data_ptr()), vp(sdf0[:, 0].data_ptr()))

        def one(a=a):
            rc = fn(pref, a[0], a[1], a[2], a[3], None, outs[0], outs[1], outs[2], outs[3], vp(torch.cuda.current_stream().cuda_stream))
            if rc != 0:
                _lib.check(rc)
        extras['step_iterate_%d_problem_iters_per_sec' % k_it] = B / (time_launches(one, 50) * 1e-3)
    _lib.set_sdf_shape(cp, IM_SIZE, IM_SIZE, IM_SIZE * IM_SIZE)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.gn_solve(cp, th0, st0, go0, sdf0, 100, 1e-4)
    barrier()
    a0.record(stream)
    sol = ops.gn_solve(cp, th0, st0, go0, sdf0, 100, 1e-4)
    a1.record(stream)
    barrier()
    n_it = float(sol[1].sum().item())
    extras['gn_solve_max_iters_100'] = {'ms': a0.elapsed_time(a1), 'mean_iters': n_it / B,
                                        'problem_iters_per_sec': n_it / (a0.elapsed_time(a1) * 1e-3)}
    cp.B = B
    clocks = sampler.stop() if sampler is not None else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- roofline ----------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (B200_PROFILING.md)'
    alg = algorithmic_bytes(B, T)
    kernel_us = ms / K * 1e3                   # this rank's average launch-to-launch duration of the one kernel in the step
    achieved = alg / (kernel_us * 1e-6) / 1e9
    traffic, prof = None, {}
    try:
        prof = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')))
        traffic = prof.get('dram_bytes_per_launch')
    except Exception:
        pass
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'kernel': 'gn_step_kernel<2,float>', 'kernel_us': kernel_us,
                'algorithmic_bytes_per_launch': alg, 'peak_source': peak_src,
                'note': 'not HBM-bound at this size: 3.2 MB per launch is 0.49 us at HBM peak; the kernel is a chain of '
                        'dependent fp64 block factorisations held in shared memory (DESIGN.md, roofline)'}
    # second view: the fp64 pipe.  Instructions per launch come from the committed ncu capture, the peak is the
    # DFMA issue rate measured on this GPU with three distinct register operands (profiles/r01_microbench.txt).
    n64, pk64 = prof.get('fp64_warp_insts_per_launch'), prof.get('dfma_warp_insts_per_cycle_per_sm_measured')
    if n64 and pk64:
        sm_hz = 1e6 * float(((clocks or {}).get('sm_mhz') or (clocks or {}).get('sm_max_mhz') or 1965.0))
        sms = torch.cuda.get_device_properties(0).multi_processor_count
        a64 = n64 / (kernel_us * 1e-6 * sm_hz * sms)
        roofline['fp64_pipe'] = {'achieved': a64, 'peak': pk64, 'unit': 'fp64 warp-instructions/cycle/SM',
                                 'frac': a64 / pk64, 'fp64_warp_insts_per_launch': n64}

    # ---------------- CPU baseline (oracle port of the reference algorithm) ----------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r, why = run_cpu_worker(6, 1, 128, 180)
        cpu = cpu_baseline_obj(r) if r is not None else {'value': None, 'unit': UNIT, 'cores': host_threads(), 'kind': 'port', 'sample': why}

    shape = ops.launch_shape(cp, torch.float32)
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': ms_max / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': '2D point robot, batch=1024 random-obstacle envs, 64 states, 1xB200 (per GPU)',
                   'batch_per_gpu': B, 'global_batch': world * B, 'states': T, 'state_dim': d, 'sdf': '%dx%d fp32' % (IM_SIZE, IM_SIZE),
                   'io_dtype': 'f32', 'iterate': ITERATE, 'parallelism': 'batch-sharded x%d, no data-path collective' % world,
                   'l2': 'rotating %d input sets (%.0f MiB) > 126 MB L2' % (N_SETS, N_SETS * (B * IM_SIZE * IM_SIZE * 4 + B * T * d * 4) / 2 ** 20),
                   'state_iters_per_sec': value * T, 'batch_iters_per_sec': value / (world * B), 'launch': shape,
                   'extras': extras,
                   'e2e_sdf_resident': {'value': e2e_res, 'unit': UNIT, 'h2d_bytes_per_step': hs.h2d_bytes,
                                        'note': 'SDF copied once and kept on the device (GN iterations on fixed environments)'}},
        'e2e': {'value': e2e_val, 'unit': UNIT, 'h2d_bytes_per_step': hs.h2d_bytes + hs.sdf_bytes,
                'd2h_bytes_per_step': hs.d2h_bytes, 'steps': Ke},
        'gpu_launches': K,
        'roofline': roofline,
        'cpu_baseline': cpu,
        'clocks': clocks,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
