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
    ap.add_argument('--quick', action='store_true',
                    help='skip the extra keys (configs[2], store ceiling)')
    ap.add_argument('--pair-mode', default=None, choices=['tma', 'pipe', 'hl', 'ab'],
                    help='pair-stage kernel (default: the engine default, tma)')
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
def cpu_threads():
    """Threads of the PyTorch-CPU restatement: every host core."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    return torch.get_num_threads()


def cpu_forward_rate(cfg, n_dets, seconds, max_images=64, first_index=0, warm=True):
    """Reference formulation on the host cores, one image per call like test.py:63-71:
    the PyTorch-CPU float32 restatement (oracle/gnet_oracle_torch.py; about 3x the numpy
    one, which spends its time in fancy indexing and reduceat), intra-op threads = all
    cores.  `warm`: one untimed forward first (callers that time several calls warm up
    once themselves)."""
    from gossipnet_b200 import params as P
    from gossipnet_b200 import synthetic
    from oracle import gnet_oracle_torch
    cpu_threads()
    layout, total = P.param_layout(1, cfg)
    net = gnet_oracle_torch.TorchGnet(P.views(layout, P.init_flat(layout, total, cfg, seed=1)),
                                      cfg, 1)
    if warm:
        net.forward(synthetic.make_image(n_dets, 1, image_index=first_index))
    done, t0 = 0, time.perf_counter()
    while done < max_images:
        net.forward(synthetic.make_image(n_dets, 1, image_index=first_index + done))
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
    per_step = 4  # images per step: a bounded sample of the workload
    for _ in range(args.warmup):
        cpu_forward_rate(cfg, args.n_dets, 1e9, max_images=1)
    t0 = time.perf_counter()
    for s in range(args.steps):
        # exactly per_step forwards per step: the warm-up forward stays outside the timer
        cpu_forward_rate(cfg, args.n_dets, 1e9, max_images=per_step, first_index=s * per_step,
                         warm=False)
    dt = time.perf_counter() - t0
    value = args.steps * per_step * args.n_dets / dt
    sample = ('%d images x N=%d per step, one image per call, %d torch intra-op threads'
              % (per_step, args.n_dets, cpu_threads()))
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': workload_name(args.n_dets, args.blocks),
                   'images_per_step': per_step},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': sample + '; PyTorch-CPU float32 restatement of the reference '
                         '(oracle/gnet_oracle_torch.py; TensorFlow 0.12 is not installable here)'},
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
def _events(torch, stream, fn, reps, before=None):
    """CUDA-event times (ms) of `fn` on `stream`, `before` (untimed) ahead of every rep."""
    out = []
    for _ in range(reps):
        if before is not None:
            before()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        stream.synchronize()
        out.append(a.elapsed_time(b))
    return out


def _load_json(name):
    try:
        return json.load(open(os.path.join(ROOT, name)))
    except Exception:
        return {}


def kernel_rooflines(torch, ops, eng, sess, stream, flush, P, T, blocks, bf16, peaks, peak_src,
                     num_classes=1):
    """The three tcgen05 kernels of the step, each timed alone with CUDA events on the
    operands the last forward left in the workspace (L2 flushed before every launch).
    A kernel timed in isolation is quoted against the BURST bf16 peak."""
    burst = float(peaks.get('bf16_tflops', 1590.0))
    ncu = _load_json('profiles/r2_ncu_traffic.json')      # dram bytes per launch from ncu --set full
    eng.want_pw_f32 = not eng._tma_path()
    res = eng.forward(sess.d_dets, sess.d_scores, sess.d_cls, sess.d_off, max_img=sess._max_img)
    eng.want_pw_f32 = True
    cap = res['capacity']
    p = eng.p
    note = ('plain bf16 operands: one tensor flop per algorithmic flop' if bf16 else
            'fp32 semantics via bf16x3: 3 tensor flops issued per algorithmic flop, so 0.333 is '
            'the ceiling of this fraction')
    src = '%s bf16 burst %.1f TF/s (kernel timed alone). %s' % (peak_src, burst, note)

    def entry(kernel, ms, flops, launches, traffic_key, extra=None):
        tf = flops / (ms * 1e-3) / 1e12
        traffic = ncu.get(traffic_key)
        if traffic is not None and 'per_pair' in str(traffic_key):
            traffic = traffic * P
        e = {'kernel': kernel, 'bound': 'tensor', 'achieved': tf, 'peak': burst, 'unit': 'TFLOP/s',
             'frac': tf / burst, 'traffic': traffic, 'ms_per_launch': ms,
             'launches_per_step': launches, 'algorithmic_flops_per_launch': flops,
             'peak_source': src}
        if extra:
            e.update(extra)
        return e

    med = lambda t: float(np.median(t[2:]))
    out = {}
    pre = lambda: flush.zero_()
    # ---- block pair stage (dominant kernel) ----------------------------------------
    pooled = eng._buf('pooled', (T, 64))
    s1 = 'gnet/block1/'
    if eng._tma_path():
        red_all = eng._ws['red_hl'][:(T + 1) * 64].view(T + 1, 64)
        ab = eng._ws['u'][:T * 64].view(T, 64)
        image, _ = eng._tma_images()
        tb = ops.pair_tma_image_bytes()

        def pair():
            ops.block_pair_fwd_tma(eng._pw_hl, red_all, T, ab, res['pair_c'], res['pair_n'],
                                   res['num_pairs'], cap, p[s1 + 'pw_fc2/biases'], image[:tb],
                                   pooled, bf16=bf16)
        name = 'block_pair_tma_kernel'
    else:
        red = eng._ws['red_hl'][:T * 64].view(T, 64)
        image, _, (pair_off, _, pair_b, _) = eng._operand_images()
        wimg = image[pair_off[0]:pair_off[0] + pair_b]

        def pair():
            ops.block_pair_fwd_pipe(res['pw_feats'], red, res['pair_c'], res['pair_n'],
                                    res['num_pairs'], cap, p[s1 + 'pw_fc1/biases'],
                                    p[s1 + 'pw_fc2/biases'], wimg, pooled, bf16=bf16)
        name = 'block_pair_pipe_kernel'
    ms = med(_events(torch, stream, pair, 10, before=lambda: (flush.zero_(), pooled.zero_())))
    out['roofline'] = entry(name, ms, 20480.0 * P, blocks, 'pair_dram_bytes_per_pair',
                            {'algorithmic_flops_per_pair': 20480,
                             'traffic_source': 'dram__bytes_read+write of one launch, ncu --set full '
                                               '(profiles/r2_ncu_traffic.json), scaled by P'})
    # ---- pair-feature MLP ------------------------------------------------------------
    if eng.fused_pw:
        w = [p[('gnet/pw_feats/fc%d/' % i) + k] for i in (1, 2, 3) for k in ('weights', 'biases')]
        hl = eng._buf('pw_hl', (cap, 64), torch.bfloat16)

        def pwfeat():
            ops.pwfeat_mlp_fwd(sess.d_dets, sess.d_scores, sess.d_cls if num_classes > 1 else None,
                               res['pair_c'], res['pair_n'], res['pair_iou'], res['num_pairs'], cap,
                               num_classes, 1.0, *w, out=None, wprep=eng._ws['wprep'], bf16=bf16,
                               out_hl=hl, want_f32=False)
        ms = med(_events(torch, stream, pwfeat, 8, before=pre))
        # 2 (w_raw 256 + 256 256 + 256 32) flop per pair, w_raw = 9 or 2C+7
        w_raw = 9 if num_classes == 1 else 2 * num_classes + 7
        out['roofline_pwfeat'] = entry('pwfeat_pipe_kernel', ms,
                                       2.0 * (w_raw * 256 + 256 * 256 + 256 * 32) * P, 1,
                                       'pwfeat_dram_bytes_per_pair')
    # ---- detection-level kernel ---------------------------------------------------------
    if eng.fused_det and blocks >= 2:
        image, _, (_, det_off, _, det_b) = eng._operand_images()
        feats = eng._buf('feats0', (T, 128))
        outb = eng._buf('feats1', (T, 128))
        inter = eng._ws['red_hl'][:T * 64].view(T, 64)
        u = eng._buf('u', (T, 64)) if eng._tma_path() else None

        def det():
            args_ = (pooled, feats, image[det_off[1]:det_off[1] + det_b], p['gnet/block1/fc1/biases'],
                     p['gnet/block1/fc2/biases'], p['gnet/block2/reduce_dim/biases'])
            if u is not None:
                fn = ops.block_det_fwd_tma if eng.det_tma else ops.block_det_fwd_img_u
                fn(*args_, outb, inter, p['gnet/block2/pw_fc1/biases'], u, bf16=bf16)
            else:
                ops.block_det_fwd_img(*args_, feats_out=outb, red_hl=inter, bf16=bf16)
        ms = med(_events(torch, stream, det, 10, before=pre))
        e = entry('block_det_tma_kernel' if (u is not None and eng.det_tma)
                  else 'block_det_tc_kernel', ms, 32768.0 * T, blocks + 1,
                  'det_dram_bytes_per_launch')
        # this kernel moves 1920 B per detection for 32 768 flop (reads: pooled 256 + shortcut
        # 512; writes: block output 512 + U 256 + red 128 + re-zeroed pooled 256): it sits on
        # the memory side of the roofline, so it is quoted against the copy bandwidth
        hbm = float(peaks.get('hbm_gbs', 6650.0))
        gbs = 1920.0 * T / (ms * 1e-3) / 1e9
        e.update({'bound': 'hbm', 'achieved': gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': gbs / hbm,
                  'algorithmic_bytes_per_launch': 1920.0 * T, 'tensor_tflops': e['achieved'],
                  'peak_source': '%s copy bandwidth (read+write); L2 flushed before the launch'
                                 % peak_src})
        out['roofline_det'] = e
    return out


def iou_sweep(torch, ops, stream, hbm_peak, peak_src, sizes=(1000, 2000, 4000, 8000, 10000)):
    """Dense det x det IoU (network.py:474-511) at the sizes of SURVEY.md 8(d): B images per
    launch so that >= 256 MB are written; algorithmic bytes 4 N^2 + 32 N per image."""
    from gossipnet_b200 import synthetic
    ncu = _load_json('profiles/r2_ncu_traffic.json').get('iou_dram_bytes', {})
    rows = []
    for n in sizes:
        b = max(1, int(np.ceil((256 << 20) / (4.0 * n * n))))
        d = torch.from_numpy(np.stack([synthetic.make_image(n, 1, image_index=i)['dets']
                                       for i in range(min(b, 4))])).cuda()
        d = d.repeat((b + d.shape[0] - 1) // d.shape[0], 1, 1)[:b].contiguous()
        out = torch.empty((b, n, n), device='cuda')

        def eight():
            for _ in range(8):
                ops.iou_dense(d, d, out=out)
        # average launch duration over 8 back-to-back launches per event pair (every launch
        # rewrites the whole >= 256 MB output, which does not fit in L2); a single launch
        # between its own event pair carries ~10 us of launch + event overhead on top
        ms = float(np.median(_events(torch, stream, eight, 6)[1:])) / 8.0
        ms1 = float(np.median(_events(torch, stream, lambda: ops.iou_dense(d, d, out=out), 6)[1:]))
        nbytes = b * (4.0 * n * n + 32.0 * n)
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append({'n': n, 'images_per_launch': b, 'bytes': nbytes, 'ms_per_launch': ms,
                     'ms_single_launch_with_event_overhead': ms1,
                     'achieved': gbs, 'frac': gbs / hbm_peak, 'traffic': ncu.get(str(n))})
        del out, d
    return rows


def store_ceiling(torch, ops, stream):
    """Sustained pure-store bandwidth of this GPU with the library's own best store kernel
    (gn_selftest_store_bw: 256-bit st.global, evict-first), 1 GiB, next to cudaMemset."""
    n = 1 << 30
    buf = torch.empty(n, dtype=torch.uint8, device='cuda')
    res = {}
    if hasattr(ops, 'selftest_store_bw'):
        for mode in ops.STORE_BW_MODES:
            t = _events(torch, stream, lambda: ops.selftest_store_bw(buf, mode), 5)
            res[mode] = n / (min(t[1:]) * 1e-3) / 1e9
    t = _events(torch, stream, lambda: buf.zero_(), 5)
    res['cudaMemset'] = n / (min(t[1:]) * 1e-3) / 1e9
    del buf
    return res


def run_config2(torch, ops, stream, flush, peaks, peak_src, steps, warmup):
    """BASELINE configs[2]: coco_multiclass (80 classes), N = 2000 detections per image,
    16 blocks, plain-bf16 operands; 16 images per step.  Not the headline: an extra key."""
    from gossipnet_b200.nms_net.config import cfg, cfg_from_file, reset_cfg
    from gossipnet_b200.nms_net.network import Gnet
    from gossipnet_b200.session import InferenceSession
    from gossipnet_b200 import synthetic
    reset_cfg()
    cfg_from_file(os.path.join(ROOT, 'experiments', 'coco_multiclass', 'conf.yaml'))
    cfg.gnet.compute_dtype = 'bf16'
    B, N, C = 16, 2000, 80
    imgs = [synthetic.make_image(N, C, seed=42, image_index=i) for i in range(B)]
    dets = np.concatenate([im['dets'] for im in imgs]).astype(np.float32)
    scores = np.concatenate([im['det_scores'] for im in imgs]).astype(np.float32)
    classes = np.concatenate([im['det_classes'] for im in imgs]).astype(np.int32)
    img_off = (np.arange(B + 1) * N).astype(np.int32)
    net = Gnet(C)
    sess = InferenceSession(net)
    sess.stream = stream
    for _ in range(max(warmup, 3)):
        sess.run(dets, scores, classes, img_off)
    P = int(sess.h_np[0])
    with torch.cuda.stream(stream):
        t = _events(torch, stream, lambda: sess._graph.replay() if sess._graph else sess._forward(sess._cur),
                    steps, before=lambda: flush.zero_())
        ms = float(np.mean(t))
        roofs = kernel_rooflines(torch, ops, net.engine, sess, stream, flush, P, B * N,
                                 cfg.gnet.num_blocks, True, peaks, peak_src, num_classes=C)
    r, rp = roofs['roofline'], roofs.get('roofline_pwfeat')
    # the C = 80 pair-feature MLP: 232 960 flop per pair (2C+7 = 167 raw features)
    reset_cfg()
    return {'workload': 'coco_multiclass 80 classes, N=2000 dets/image, 16 blocks, plain bf16 '
                        'operands + fp32 accumulation (BASELINE configs[2]); %d images per step' % B,
            'value': B * N / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'steps': steps,
            'pairs_per_step': P, 'dtype': 'bf16',
            'roofline': {k: r[k] for k in ('kernel', 'achieved', 'peak', 'unit', 'frac',
                                           'ms_per_launch', 'peak_source')},
            'roofline_pwfeat': None if rp is None else
            {k: rp[k] for k in ('kernel', 'achieved', 'peak', 'unit', 'frac', 'ms_per_launch')}}


def run_train(torch, dist, world, rank, steps, warmup):
    """BASELINE configs[3]: training, 8 images x N=1000 per GPU (64 images on 8 GPUs), 16 blocks:
    forward with kept activations + DetectionMatching + loss + backward + ONE NCCL all-reduce of
    the flat gradient buffer + fused Adam (train.py:64-77, 316-320).  Device time per step, max
    over ranks; the all-reduce also timed alone."""
    from gossipnet_b200 import synthetic
    from gossipnet_b200.nms_net.network import Gnet
    from gossipnet_b200.trainer import Trainer
    per_gpu, n_dets = 8, 1000
    imgs = [synthetic.make_image(n_dets, 1, image_index=rank * per_gpu + i) for i in range(per_gpu)]
    net = Gnet(1)
    tr = Trainer(net)
    for _ in range(max(2, warmup)):
        tr.step(imgs, 1e-4)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        res = tr.step(imgs, 1e-4)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    # the collective alone: the flat fp32 gradient buffer (+ the image count)
    ar = []
    for _ in range(12):
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if world > 1:
            dist.all_reduce(tr.gradbuf)
        e1.record()
        torch.cuda.synchronize()
        ar.append(e0.elapsed_time(e1) * 1e3)
    ar_us = float(np.median(ar[2:]))
    # replicas must stay bit-identical: same all-reduced gradient, same update on every rank
    same = True
    t = torch.tensor([ms, ar_us], dtype=torch.float64, device='cuda')
    if world > 1:
        ref = tr.eng.flat.clone()
        dist.broadcast(ref, 0)
        diff = (ref != tr.eng.flat).any().to(torch.int32)
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        same = int(diff) == 0
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ar_us = float(t[0]), float(t[1])
    return {'workload': 'training step, %d images x N=%d per GPU, 16 blocks, coco_person '
                        '(BASELINE configs[3]: batch 64 on 8 GPUs); forward with kept activations, '
                        'matching, loss, backward, all-reduce, Adam' % (per_gpu, n_dets),
            'ms_per_step': ms, 'value': world * per_gpu * n_dets / (ms * 1e-3), 'unit': UNIT,
            'steps': steps, 'pairs_per_step_rank0': int(res['P']),
            'allreduce_us': ar_us if world > 1 else None,
            'allreduce_bytes': int(tr.gradbuf.numel() * 4),
            'params_identical_across_ranks': same,
            'arithmetic': 'fp32 semantics: every FC and its gradients as bf16x3 tcgen05 GEMMs '
                          '(gn_fc_tc.cu)'}


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
    peaks = _load_json('MEASURED_PEAKS.json')
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured' if peaks else 'fallback'

    B, N = args.images, args.n_dets
    imgs, dets, scores, classes, img_off = make_inputs(B, N, first_index=rank * B)
    net = Gnet(1)
    eng = net.engine
    if args.pair_mode:
        eng.pair_mode = args.pair_mode
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
        # (a) the streaming call a batch loop makes (test.py): one batch ahead, every step's
        #     inputs pinned host -> device and its logits device -> pinned host inside the
        #     timed region; (b) one synchronous call per step
        one = (dets, scores, classes, img_off)
        for pred in sess.run_pipelined([one] * 4):      # captures the two pipeline slots
            pass
        barrier()
        t0 = time.perf_counter()
        chk = 0.0
        for pred in sess.run_pipelined([one] * args.steps):
            chk += float(pred[0])                         # the result is read on the host
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        barrier()
        t0 = time.perf_counter()
        for s in range(args.steps):
            pred = sess.run(dets, scores, classes, img_off)
        torch.cuda.synchronize()
        e2e_sync_s = time.perf_counter() - t0
    clk = clocks.summary()
    t = torch.tensor([dev_ms, e2e_s * 1e3, e2e_sync_s * 1e3], dtype=torch.float64, device='cuda')
    pstat = torch.tensor([float(P), -float(P)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(pstat, op=dist.ReduceOp.MAX)    # max P, -min P over the ranks' shards
    dev_ms, e2e_ms, e2e_sync_ms = float(t[0]), float(t[1]), float(t[2])
    p_max, p_min = int(pstat[0]), int(-pstat[1])
    total_dets = world * B * N * args.steps
    value = total_dets / (dev_ms * 1e-3)
    e2e_value = total_dets / (e2e_ms * 1e-3)

    # ---- rooflines: the tensor-core kernels timed live (un-graphed, events per launch) --
    roofs, sweep, extra2, stores = {}, None, None, None
    if rank == 0:
        with torch.cuda.stream(stream):
            roofs = kernel_rooflines(torch, ops, eng, sess, stream, flush, P, T, args.blocks, bf16,
                                     peaks, peak_src)
            sweep = iou_sweep(torch, ops, stream, hbm_peak, peak_src)
            if world == 1 and not args.quick:
                stores = store_ceiling(torch, ops, stream)
    roof_iou = None
    if sweep:
        big = sweep[-1]
        roof_iou = {'kernel': 'iou_symmetric_kernel', 'bound': 'hbm', 'achieved': big['achieved'],
                    'peak': hbm_peak, 'unit': 'GB/s', 'frac': big['frac'], 'traffic': big['traffic'],
                    'ms_per_launch': big['ms_per_launch'],
                    'workload': 'N=M=%d, one image' % big['n'],
                    'peak_source': peak_src + ' copy bandwidth (read+write)',
                    'sweep': sweep, 'store_ceiling_gbs': stores}

    # ---- CPU baseline on this box's cores (rank 0, N=1) ---------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, n_img, dt = cpu_forward_rate(cfg, N, args.cpu_seconds)
        cpu = {'value': rate, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
               'sample': '%d images x N=%d, one image per call, %.1f s, %d torch intra-op threads; '
                         'PyTorch-CPU float32 restatement of the reference '
                         '(oracle/gnet_oracle_torch.py)' % (n_img, N, dt, cpu_threads())}
    # ---- BASELINE configs[2] (extra key; single GPU, default run) -------------------
    if rank == 0 and world == 1 and not args.quick and not bf16:
        extra2 = run_config2(torch, ops, stream, flush, peaks, peak_src, max(5, args.steps // 2),
                             args.warmup)

    # ---- BASELINE configs[3]: the training step and its one collective, at every N -------
    train = None
    pair_mode, used_graph = eng.pair_mode, bool(sess._graph)
    h2d_bytes, d2h_bytes = sess.h2d_bytes, sess.d2h_bytes
    if not args.quick and not bf16:
        del sess, net, eng, flush
        torch.cuda.empty_cache()
        setup_cfg(args.blocks, args.precision)
        train = run_train(torch, dist, world, rank, 5, 2)

    if rank == 0:
        line = {
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
                       'pairs_per_step_min_max_over_ranks': [p_min, p_max],
                       'pair_mode': pair_mode,
                       'parallelism': 'images sharded over %d GPU(s), no data-path collective'
                                      % world,
                       'l2': 'flushed between timed steps (256 MiB memset); working set also '
                             '> 126 MB L2',
                       'cuda_graph': used_graph},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d_bytes,
                    'd2h_bytes_per_step': d2h_bytes, 'ms_per_step': e2e_ms / args.steps,
                    'api': 'gossipnet_b200.session.InferenceSession.run_pipelined (numpy in/out, '
                           'one batch ahead; the loop test.py runs)',
                    'sync_call': {'value': total_dets / (e2e_sync_ms * 1e-3),
                                  'ms_per_step': e2e_sync_ms / args.steps,
                                  'api': 'InferenceSession.run, one blocking call per step'}},
            'gpu_launches': launches * args.steps,
            'clocks': clk, 'roofline': roofs.get('roofline'),
            'roofline_pwfeat': roofs.get('roofline_pwfeat'), 'roofline_det': roofs.get('roofline_det'),
            'roofline_iou': roof_iou, 'cpu_baseline': cpu, 'config2_coco_multiclass_bf16': extra2,
            'train_config3': train,
            'logit_checksum': float(np.sum(pred, dtype=np.float64)),
        }
        print(json.dumps(line))
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
