#!/usr/bin/env python
"""bench.py: detections/s of the Gnet forward (N=1000 detections/image, 16 blocks,
coco_person hyper-parameters = BASELINE.json configs[1]) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One step = one forward of the hot path (neighbor build -> pair-feature MLP ->
16 blocks -> predict head) over a batch of `--images` synthetic images per GPU.
Images shard across ranks with no data-path collective (weak scaling).

Printed JSON (one line, rank 0): `value` = whole-job detections/s with inputs
resident in HBM (CUDA-graph replay, CUDA events, L2 flushed between steps, max
over ranks); `e2e` = the same metric through the host-buffer session
(gossipnet_b200.session.InferenceSession.run: pinned H2D of the step's inputs
and D2H of its logits inside the timed region); `roofline` = the dominant
kernel (block pair stage) timed live with CUDA events; `roofline_iou` = the
dense N x N IoU kernel (BASELINE metric "IoU HBM GB/s"); `cpu_baseline` = the
numpy restatement of the reference (oracle/, TensorFlow is not installable)
timed on this box's host cores on a bounded sample.

`--impl reference` times that CPU restatement alone (rank 0 only).
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'detections/sec Gnet fwd (N=1000, 16 blocks)'
UNIT = 'detections/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--images', type=int, default=64, help='images per GPU per step')
    ap.add_argument('--n-dets', type=int, default=1000)
    ap.add_argument('--blocks', type=int, default=16)
    ap.add_argument('--cpu-seconds', type=float, default=12.0,
                    help='budget of the cpu_baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--precision', default='fp32', choices=['fp32', 'bf16'],
                    help="arithmetic of the fused FC kernels: fp32 semantics (bf16x3, the "
                         "headline) or plain bf16 operands (BASELINE configs[2]'s arithmetic)")
    return ap.parse_args()


def workload_name(n_dets, blocks):
    return ('coco_person N=%d dets/image, %d blocks, d=128, fp32 (BASELINE configs[1])'
            % (n_dets, blocks))


def setup_cfg(blocks, precision='fp32'):
    from gossipnet_b200.nms_net.config import cfg, cfg_from_file, reset_cfg
    reset_cfg()
    cfg_from_file(os.path.join(ROOT, 'experiments', 'coco_person', 'conf.yaml'))
    cfg.gnet.num_blocks = blocks
    cfg.gnet.compute_dtype = precision
    return cfg


def make_inputs(n_images, n_dets, first_index):
    from gossipnet_b200 import synthetic
    imgs = [synthetic.make_image(n_dets, 1, seed=42, image_index=first_index + i)
            for i in range(n_images)]
    dets = np.concatenate([im['dets'] for im in imgs]).astype(np.float32)
    scores = np.concatenate([im['det_scores'] for im in imgs]).astype(np.float32)
    classes = np.concatenate([im['det_classes'] for im in imgs]).astype(np.int32)
    img_off = (np.arange(n_images + 1) * n_dets).astype(np.int32)
    return imgs, dets, scores, classes, img_off


# ------------------------------------------------------------------ CPU baseline
def blas_threads():
    """Threads the BLAS behind numpy uses (the FCs of the CPU restatement run there)."""
    try:
        from threadpoolctl import threadpool_info
        n = [int(i.get('num_threads', 0)) for i in threadpool_info() if i.get('user_api') == 'blas']
        return max(n) if n else None
    except Exception:
        return None


def cpu_forward_rate(cfg, n_dets, seconds, max_images=64, first_index=0, warm=True):
    """Reference formulation on the host cores, one image per call like
    test.py:63-71 (numpy float32 restatement, BLAS threads = all cores).  `warm`: one
    untimed forward first (callers that time several calls warm up once themselves)."""
    from gossipnet_b200 import params as P
    from gossipnet_b200 import synthetic
    from oracle import gnet_oracle
    layout, total = P.param_layout(1, cfg)
    pv = P.views(layout, P.init_flat(layout, total, cfg, seed=1))
    keys = ('dets', 'det_scores', 'det_classes')
    if warm:
        img = synthetic.make_image(n_dets, 1, image_index=first_index)
        gnet_oracle.gnet_forward({k: img[k] for k in keys}, pv, cfg, 1, keep_intermediates=False)
    done, t0 = 0, time.perf_counter()
    while done < max_images:
        img = synthetic.make_image(n_dets, 1, image_index=first_index + done)
        gnet_oracle.gnet_forward({k: img[k] for k in keys}, pv, cfg, 1, keep_intermediates=False)
        done += 1
        if time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    return done * n_dets / dt, done, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = setup_cfg(args.blocks)
    cores = os.cpu_count()
    per_step = 2  # images per step: a bounded sample of the workload
    for _ in range(args.warmup):
        cpu_forward_rate(cfg, args.n_dets, 1e9, max_images=1)
    t0 = time.perf_counter()
    for s in range(args.steps):
        # exactly per_step forwards per step: the warm-up forward stays outside the timer
        cpu_forward_rate(cfg, args.n_dets, 1e9, max_images=per_step, first_index=s * per_step,
                         warm=False)
    dt = time.perf_counter() - t0
    value = args.steps * per_step * args.n_dets / dt
    sample = ('%d images x N=%d per step, one image per call, %s BLAS threads'
              % (per_step, args.n_dets, blas_threads()))
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': workload_name(args.n_dets, args.blocks),
                   'images_per_step': per_step},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': sample + '; numpy float32 restatement of the reference '
                         '(oracle/gnet_oracle.py; TensorFlow 0.12 is not installable here)'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


# ------------------------------------------------------------------------ clocks
class ClockSampler(object):
    """SM clock + throttle reasons DURING the timed region.  NVML is polled from a thread
    every 10 ms (a timed region of K x 6 ms is shorter than nvidia-smi's start-up, which left
    short multi-GPU runs without a single sample); `nvidia-smi -lms` is the fallback."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')
    NAMES = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
    BITS = (0x8, 0x40, 0x20, 0x4)      # nvmlClocksEventReason{HwSlowdown,HwThermalSlowdown,SwThermalSlowdown,SwPowerCap}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml = self._physical_index(index), [], None, None
        self._stop = threading.Event()

    @staticmethod
    def _physical_index(local):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
        parts = [p.strip() for p in vis.split(',') if p.strip()]
        if local < len(parts) and parts[local].isdigit():
            return int(parts[local])
        return local

    def _poll_nvml(self):
        n, h = self.nvml
        while not self._stop.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
                try:
                    reasons = n.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    reasons = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(sm), self._max, int(reasons)))
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        try:
            import pynvml as n
            n.nvmlInit()
            h = n.nvmlDeviceGetHandleByIndex(self.index)
            self._max = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
            self.nvml = (n, h)
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(',')]
            try:
                bits = sum(b for b, v in zip(self.BITS, r[2:6]) if v.lower().startswith('active'))
                self.rows.append((float(r[0]), float(r[1]), bits))
            except Exception:
                continue

    def __exit__(self, *a):
        self._stop.set()
        if self.nvml is not None:
            self.thread.join(timeout=1)
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        reasons = set()
        for _, _, bits in self.rows:
            for name, b in zip(self.NAMES, self.BITS):
                if bits & b:
                    reasons.add(name)
        return {'sm_mhz': float(np.median([r[0] for r in self.rows])),
                'sm_max_mhz': float(np.max([r[1] for r in self.rows])),
                'reasons': sorted(reasons), 'samples': len(self.rows),
                'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}


# ------------------------------------------------------------------------- B200
def run_b200(args):
    import torch
    import torch.distributed as dist
    from gossipnet_b200 import _lib, ops
    from gossipnet_b200.nms_net.network import Gnet
    from gossipnet_b200.session import InferenceSession

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner to stdout at the
        # VERSION / WARN debug levels, and honours NCCL_DEBUG_FILE only above VERSION
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'WARN'
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    cfg = setup_cfg(args.blocks, args.precision)
    bf16 = args.precision == 'bf16'
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    bf16_peak = float(peaks.get('bf16_tflops_sustained', 1400.0))
    peak_src = 'measured' if peaks else 'fallback'

    B, N = args.images, args.n_dets
    imgs, dets, scores, classes, img_off = make_inputs(B, N, first_index=rank * B)
    net = Gnet(1)
    eng = net.engine
    sess = InferenceSession(net, use_graph=not args.no_graph)
    T = dets.shape[0]

    # ---- e2e warm-up doubles as graph capture ------------------------------------
    for _ in range(max(args.warmup, 3)):
        pred = sess.run(dets, scores, classes, img_off)
    P = int(sess.h_np[0])
    launches = sess.launches_per_forward
    stream = sess.stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def replay():
        if sess._graph:
            sess._graph.replay()
        else:
            sess._forward(sess._cur)

    # ---- value: inputs resident, device time, L2 flushed between steps ----------
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            replay()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    with ClockSampler(local) as clocks:
        with torch.cuda.stream(stream):
            for s in range(args.steps):
                flush.zero_()
                ev[s][0].record(stream)
                replay()
                ev[s][1].record(stream)
        barrier()
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        # ---- e2e: host buffers in, host logits out ------------------------------
        barrier()
        t0 = time.perf_counter()
        for s in range(args.steps):
            pred = sess.run(dets, scores, classes, img_off)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
    clk = clocks.summary()
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    total_dets = world * B * N * args.steps
    value = total_dets / (dev_ms * 1e-3)
    e2e_value = total_dets / (e2e_ms * 1e-3)

    # ---- roofline: the dominant kernel, timed live (un-graphed, events per launch)
    roof = roof_iou = None
    if rank == 0:
        with torch.cuda.stream(stream):
            res = eng.forward(sess.d_dets, sess.d_scores, sess.d_cls, sess.d_off)
            # the operands exactly as the forward leaves them: bf16 (hi | lo) reduced
            # features of the last block, fp32 pw_feats, the pair lists, block 1's operand image
            red = eng._ws['red_hl'][:T * 64].view(T, 64)
            pooled = eng._buf('pooled', (T, 64))
            s1 = 'gnet/block1/'
            image, _, (pair_off, _, pair_b, _) = eng._operand_images()
            wimg = image[pair_off[0]:pair_off[0] + pair_b]
            times = []
            for rep in range(8):
                flush.zero_()
                pooled.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                ops.block_pair_fwd_pipe(res['pw_feats'], red, res['pair_c'], res['pair_n'],
                                        res['num_pairs'], res['capacity'],
                                        eng.p[s1 + 'pw_fc1/biases'], eng.p[s1 + 'pw_fc2/biases'],
                                        wimg, pooled, bf16=bf16)
                b.record(stream)
                stream.synchronize()
                times.append(a.elapsed_time(b))
            k_ms = float(np.median(times[2:]))
        flops = 20480.0 * P  # 2*(96*64 + 64*64) per pair
        roof = {'kernel': 'block_pair_pipe_kernel', 'bound': 'tensor',
                'achieved': flops / (k_ms * 1e-3) / 1e12, 'peak': bf16_peak, 'unit': 'TFLOP/s',
                'frac': flops / (k_ms * 1e-3) / 1e12 / bf16_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel on
                # this workload (ncu --set full, profiles/r1_tc_kernels.md); scales with P
                'traffic': 378.9e6 * (P / 2486884.0),
                'ms_per_launch': k_ms, 'launches_per_step': args.blocks,
                'algorithmic_flops_per_launch': flops,
                'peak_source': '%s bf16 sustained %.1f TF/s (kernel timed inside a long step). %s'
                               % (peak_src, bf16_peak,
                                  'plain bf16 operands: one tensor flop per algorithmic flop' if bf16
                                  else 'fp32 semantics via bf16x3: the kernel issues 3 tensor flops '
                                       'per algorithmic flop, so 0.333 is the ceiling of this fraction')}
        # dense IoU kernel at the stress size (N=10000 -> 400 MB written, > L2)
        with torch.cuda.stream(stream):
            n_iou = 10000
            from gossipnet_b200 import synthetic
            big = torch.from_numpy(synthetic.make_image(n_iou, 1)['dets']).cuda()
            out = torch.empty((1, n_iou, n_iou), device='cuda')
            times = []
            for rep in range(8):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                ops.iou_dense(big.unsqueeze(0), big.unsqueeze(0), out=out)
                b.record(stream)
                stream.synchronize()
                times.append(a.elapsed_time(b))
            i_ms = float(np.median(times[2:]))
            del out
            # context for a write-only kernel: sustained pure-store bandwidth of this GPU
            # (2 GiB memset, >> L2), next to the copy peak the fraction is quoted against
            wbuf = torch.empty(1 << 31, dtype=torch.uint8, device='cuda')
            wt = []
            for rep in range(4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                wbuf.zero_()
                b.record(stream)
                stream.synchronize()
                wt.append(a.elapsed_time(b))
            write_peak = (1 << 31) / (min(wt[1:]) * 1e-3) / 1e9
            del wbuf
        iou_bytes = 4.0 * n_iou * n_iou + 16.0 * 2 * n_iou
        roof_iou = {'kernel': 'iou_symmetric_kernel', 'bound': 'hbm',
                    'achieved': iou_bytes / (i_ms * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                    'frac': iou_bytes / (i_ms * 1e-3) / 1e9 / hbm_peak, 'traffic': None,
                    'ms_per_launch': i_ms, 'workload': 'N=M=%d, one image' % n_iou,
                    'peak_source': peak_src + ' copy bandwidth (read+write)',
                    'write_only_peak_measured_here': write_peak,
                    'frac_of_write_only_peak': iou_bytes / (i_ms * 1e-3) / 1e9 / write_peak}

    # ---- CPU baseline on this box's cores (rank 0, N=1) ---------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, n_img, dt = cpu_forward_rate(cfg, N, args.cpu_seconds)
        cpu = {'value': rate, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
               'sample': '%d images x N=%d, one image per call, %.1f s; numpy float32 restatement '
                         'of the reference (oracle/gnet_oracle.py)' % (n_img, N, dt)}

    if rank == 0:
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dev_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16' if bf16 else 'f32',
            'data': 'synthetic',
            'config': {'workload': workload_name(N, args.blocks),
                       'arithmetic': ('fp32 in/out; FC GEMMs with plain bf16 operands on tcgen05, '
                                      'fp32 TMEM accumulation (--precision bf16; NOT the headline)'
                                      if bf16 else
                                      'fp32 in/out; FC GEMMs as bf16 hi/lo split products '
                                      '(bf16x3) on tcgen05 with fp32 TMEM accumulation'),
                       'images_per_gpu_per_step': B, 'pairs_per_step_rank0': P,
                       'parallelism': 'images sharded over %d GPU(s), no data-path collective'
                                      % world,
                       'l2': 'flushed between timed steps (256 MiB memset); working set also '
                             '> 126 MB L2',
                       'cuda_graph': bool(sess._graph)},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': sess.h2d_bytes,
                    'd2h_bytes_per_step': sess.d2h_bytes, 'ms_per_step': e2e_ms / args.steps,
                    'api': 'gossipnet_b200.session.InferenceSession.run (numpy in/out)'},
            'gpu_launches': launches * args.steps,
            'clocks': clk, 'roofline': roof, 'roofline_iou': roof_iou, 'cpu_baseline': cpu,
            'logit_checksum': float(np.sum(pred, dtype=np.float64)),
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
