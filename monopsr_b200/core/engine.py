"""B200 execution engine of the MonoPSR per-instance network (forward, backward, train-op).

Host-side orchestration only: every arithmetic operation is a hand-written sm_100a kernel
reached through the C ABI of ``include/monopsr_b200_net.h``; torch tensors are device-memory
containers (plus ``torch.distributed`` for the one gradient all-reduce).  The op sequence
restates the reference graph:

  builders/net_builder.py:30-96                         two ResNet-101 towers, squash, map decoder
  core/models/monopsr/monopsr_output_builder.py          FC stacks and heads
  core/models/monopsr/monopsr_model.py:138-492,554-958   wiring and losses
  core/trainer.py:76-81, builders/optimizer_builder.py   train-op

One ``Engine`` = one sample in flight (num_boxes crops + one full image), matching the
reference's ``batch_size: 1`` (configs/monopsr_model_000.yaml:14-17).  The whole step is
captured into a CUDA graph after the first call.
"""
import ctypes
import os
import math

import numpy as np
import torch

from .. import lib as _lib
from ..lib_net import TC_DGRAD, TC_FWD, TC_WGRAD, BnLayer, HeadsIO, OptChunk, TcGemmParams, W16Layer
from . import model_spec as ms

BN_EPS_RESNET = 1e-5
BN_EPS_DECODER = 1e-3
BN_DECAY_DECODER = 0.999
# train-op constants (optimizer_builder.py:23-118 with monopsr_model_000.yaml:141-152; trainer.py:76-81).  Pinned to
# the calls the reference's own optimizer_builder makes: tests/test_arch_golden.py::test_train_op_constants
LR_INITIAL, LR_DECAY_STEPS, LR_DECAY_FACTOR = 0.00008, 10000, 0.8          # exponential_decay, staircase
ADAM_BETA1, ADAM_BETA2, ADAM_EPSILON = 0.9, 0.999, 1e-8                   # tf.train.AdamOptimizer defaults
EMA_DECAY = 0.9999                                                        # MovingAverageOptimizer(average_decay)
CLIP_GRADIENT_NORM = 1.0                                                  # per variable (slim.learning.create_train_op)
KPAD = 1088          # 1043 / 1060 concat widths padded to a multiple of 64
DEFAULT_PRECISION = "h3"


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class Engine:
    _rounding_now = {}          # device index -> operand-rounding state of the library (starts at 1)
    # train-op hyper-parameters: class defaults = monopsr_model_000.yaml's; configure() overrides them per engine
    lr_initial, lr_decay_steps, lr_decay_factor, lr_staircase, ema_decay = LR_INITIAL, LR_DECAY_STEPS, LR_DECAY_FACTOR, True, EMA_DECAY

    def __init__(self, device, params=None, num_boxes=ms.NUM_BOXES, seed=0, sms=148, precision=None,
                 xyz_loss=("smooth_l1_nonzero", 100.0)):
        self.dev = torch.device(device)
        self.N = num_boxes
        self.L = _lib.load()
        self.sms = sms
        self._launch_checks = True
        # Forward precision -- all three keep fp32 accumulation in TMEM:
        #   "h3"   forward GEMMs on fp16 hi/lo SPLIT COPIES of both operands, three products as six kind::f16 MMAs per
        #          k-block (csrc/tc_gemm.cu H3 branch, csrc/split16.cu): fp32-level accuracy -- every forward output
        #          within 1e-3 of the fp32 reference -- for 1.5x the tensor work of one tf32 pass at unchanged operand
        #          traffic.  fp16's range is watched: Engine.check_overflow() raises if an activation beyond 65504 was split.
        #   "x3"   3xTF32 on unrounded operands, lo tiles derived on chip (tc_gemm_x3_kernel): same accuracy, 3x the
        #          forward tensor work, no range limit -- the fallback when h3 overflows.
        #   "tf32" single pass on operands rounded to nearest at their producers: the documented FAST mode, misses the
        #          1e-3 bar on the decoder maps (2.6e-3).
        # The backward pass is single-pass tf32 in every mode.  The operand-rounding switch is a __constant__ of the
        # library: _enter() re-asserts it whenever engines of different precision alternate inside one process.
        self.precision = (precision or os.environ.get("MPB_PRECISION", DEFAULT_PRECISION)).lower()
        if self.precision not in ("h3", "x3", "tf32"):
            raise ValueError("precision must be 'h3', 'x3' or 'tf32' (got %r)" % self.precision)
        self.x3 = self.precision == "x3"
        self.h3 = self.precision == "h3"
        # x3 / h3 read their forward operands unrounded (h3: the split copies are taken from the unrounded values), so
        # every producer leaves its result as computed; the single-pass backward then sees fp32 words whose low 13 bits
        # the tensor core drops (truncation; measured gradient error in these modes: median 4e-3, below tf32 mode's)
        self.rounding = 1 if self.precision == "tf32" else 0
        self._s16 = {}              # storage pointer -> split copy (h3)
        # loss of the local xyz map: yaml loss_config.inst_xyz_map_local = [type, weight] (loss_builder.py:19-84,
        # monopsr_model.py:580-586).  'smooth_l1_nonzero' is fused into the heads kernel; 'chamfer_dist' / 'emd' run
        # the point-set ops inside the step (value AND gradient: nn_distance_grad / matchcostgrad feed d_xyz).
        self.xyz_loss_type, self.xyz_loss_weight = str(xyz_loss[0]), float(xyz_loss[1])
        if self.xyz_loss_type not in ("smooth_l1_nonzero", "chamfer_dist", "emd"):
            raise NotImplementedError("loss type %r for inst_xyz_map_local has no sm_100a kernel" % self.xyz_loss_type)
        self._ps = None             # buffers of the point-set loss, allocated on first use
        self._enter()
        with torch.cuda.device(self.dev):
            self._build_param_layout()
            self._alloc_state()
            self.load_params(params if params is not None else ms.init_params(seed))
            self._alloc_activations()
        self.step_count = 0
        self.world = 1
        # side streams: the two towers are independent until the concat, and weight gradients are
        # off the critical path of the data-gradient chain -> they run concurrently (also inside
        # the captured CUDA graph, where the fork/join become graph edges)
        # Priorities (recorded per kernel node at capture): the block scheduler dispatches queued grids in
        # submission order, so a short kernel of a latency-critical chain can wait tens of microseconds behind the
        # CTAs of a bulk launch from another stream.  The dependent chains (main, full-image tower, FC stacks) run
        # at high priority; weight gradients, zero-fills and the side train-op are filler at default priority.
        # Measured: 9.09 vs 8.55 ms/step -- the starved weight-gradient streams finish late -- so OFF by default.
        # MPB_STREAM_PRIO=2: only the FC stacks' stream -- a chain of SMALL kernels whose backward half gates the towers'
        # backward pass and which the timeline shows waiting 20-100 us at a time behind the decoder's bulk weight gradients
        prio = int(os.environ.get("MPB_STREAM_PRIO", "0"))
        hi = -1 if prio == 1 else 0
        self.s_main = torch.cuda.Stream(device=self.dev, priority=hi)     # capture stream of the step
        self.s_full = torch.cuda.Stream(device=self.dev, priority=hi)
        self.s_fc = torch.cuda.Stream(device=self.dev, priority=-1 if prio in (1, 2) else 0)
        self.s_wc = torch.cuda.Stream(device=self.dev)
        self.s_wf = torch.cuda.Stream(device=self.dev)
        self.s_opt = torch.cuda.Stream(device=self.dev)
        self.s_ar = torch.cuda.Stream(device=self.dev)      # d(gamma) + all-reduce issue of the gradient buckets
        self.early_opt = False      # set by the single-GPU train step: head train-op under the towers' backward
        self.early_opt_unit = None  # ... started when the full-image tower's backward chain reaches this unit
        self._grads_zeroed = False
        self.overlap = True
        # tile planning knobs (tools/r2_call_f.sh / r2_call_h.sh sweeps, profiles/r2_notes.md).  The h3 forward does 1.5x
        # the MMA work per operand byte, which moves the balance towards wider tiles: 128-wide tiles already at 0.6 of
        # the SMs and for the short epilogue-bound reductions measured 9.11 vs 9.40 ms/step (tf32 mode keeps 0.9 / 64)
        wide = self.precision != "tf32"
        self.fill = float(os.environ.get("MPB_TILE_FILL", "0.6" if wide else "0.9"))   # min fraction of SMs a launch must fill before widening tiles
        self.bn_fused = int(os.environ.get("MPB_BN_FUSED", "1")) != 0      # decoder batch norm: one launch per direction
        self.shortk = int(os.environ.get("MPB_SHORTK", "512"))
        self.shortk_bn = int(os.environ.get("MPB_SHORTK_BN", "128" if wide else "64"))   # tile width of short, epilogue-bound reductions
        self.wgrad_bn = int(os.environ.get("MPB_WGRAD_BN", "128"))
        self.wgrad_fill = float(os.environ.get("MPB_WGRAD_FILL", "0.5"))   # target CTAs / SMs when choosing split-K
        # cluster split-K (DSMEM reduce) for long reductions: faster per launch (profiles/r1_gemm_sweep.txt), but the
        # step is 2.5 % faster WITHOUT it once launches are chained with programmatic dependent launch (cluster
        # launches do not overlap their predecessor's tail): off by default, MPB_CSK=1 enables
        self.csk = int(os.environ.get("MPB_CSK", "0"))
        self.csk_fwd = int(os.environ.get("MPB_CSK_FWD", str(self.csk)))   # the forward pass alone (only 2 streams there)
        self.csk_bn = int(os.environ.get("MPB_CSK_BN", "128"))            # widest tile that may be split over a cluster
        self.ctas_per_sm = {64: 2, 128: int(os.environ.get("MPB_CTAS128", "2")), 256: 1}   # see tc_gemm.cuh

    # ------------------------------------------------------------------ parameters
    def _dev_shape(self, name, shape, kind):
        if kind == "weights":
            if len(shape) == 4:
                k, _, cin, cout = shape
                return (cout, k * k * cin)
            kin, kout = shape
            if kin in (1043, 1060):
                kin = KPAD
            return (kout, kin)
        return tuple(shape)

    def _build_param_layout(self):
        table = ms.param_table()
        self.ptable = {n: (s, k) for n, s, k in table}
        self.layout = {}          # name -> (arena, offset, dev_shape)
        off = 0
        self.trainable_names = []
        for n, s, k in table:
            if k in ms.TRAINABLE_KINDS:
                ds = self._dev_shape(n, s, k)
                self.layout[n] = ("T", off, ds)
                off += int(np.prod(ds))
                off = (off + 3) & ~3          # 16-byte alignment of every tensor
                self.trainable_names.append(n)
        self.n_train = off
        off = 0
        for n, s, k in table:
            if k not in ms.TRAINABLE_KINDS:
                self.layout[n] = ("S", off, tuple(s))
                off += int(np.prod(s))
                off = (off + 3) & ~3
        self.n_state = off

    def _alloc_state(self):
        z = lambda n, dt=torch.float32: torch.zeros(n, dtype=dt, device=self.dev)
        self.params = z(self.n_train)
        self.state = z(self.n_state)
        self.grads = z(self.n_train)
        self.adam_m = z(self.n_train)
        self.adam_v = z(self.n_train)
        self.ema = z(self.n_train)
        self.prep = z(self.n_train)            # folded / tf32-rounded weights (same offsets as params)
        self.prep16 = z(self.n_train) if self.h3 else None     # [lo | hi] fp16 split copies of the forward GEMM weights
        self.overflow = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.hyper = z(4)
        # optimizer chunk table
        chunks = []
        CH = 1 << 16
        for ti, n in enumerate(self.trainable_names):
            _, off, ds = self.layout[n]
            size = int(np.prod(ds))
            for s in range(0, size, CH):
                chunks.append((off + s, min(CH, size - s), ti))
        arr = (OptChunk * len(chunks))()
        for i, (a, b, c) in enumerate(chunks):
            arr[i].start, arr[i].len, arr[i].tensor = a, b, c
        raw = bytes(arr)
        self.opt_chunks = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.dev)
        self.n_chunks = len(chunks)
        self.norm2 = z(len(chunks))            # per-chunk sums of squares (combined per variable in a fixed order)
        # the variables outside the two towers (squash, decoder, FC stacks, heads: 45 % of the parameters) sit at
        # the end of the arena; their gradients are final before the towers' backward pass starts, so their
        # train-op runs on a side stream under it ("head" part; the rest is the "towers" part)
        t_head = min(i for i, n in enumerate(self.trainable_names) if not n.startswith("FirstStage"))
        assert all(not n.startswith("FirstStage") for n in self.trainable_names[t_head:])
        c_head = min(i for i, c in enumerate(chunks) if c[2] >= t_head)
        nt = len(self.trainable_names)
        self.opt_parts = {"all": (0, len(chunks), 0, nt), "towers": (0, c_head, 0, t_head),
                          "head": (c_head, len(chunks) - c_head, t_head, nt - t_head)}
        # the towers' part again, one tower each (data parallelism: the second tower's gradients are still on NVLink
        # while the first tower's variables are stepped)
        second = self.trainable_names[0].split("/")[0]
        t_mid = min(i for i, n in enumerate(self.trainable_names[:t_head]) if n.split("/")[0] != second)
        assert all(n.split("/")[0] == second for n in self.trainable_names[:t_mid])
        assert len({n.split("/")[0] for n in self.trainable_names[t_mid:t_head]}) == 1
        c_mid = min(i for i, c in enumerate(chunks) if c[2] >= t_mid)
        self.tower_mid = chunks[c_mid][0]              # arena offset where the second tower's variables start
        assert all(c[0] + c[1] <= self.tower_mid for c in chunks[:c_mid])
        self.opt_parts["tower0"] = (0, c_mid, 0, t_mid)
        self.opt_parts["tower1"] = (c_mid, c_head - c_mid, t_mid, t_head - t_mid)

    def view(self, name, arena=None):
        a, off, ds = self.layout[name]
        base = {"T": self.params, "S": self.state}[a] if arena is None else arena
        return base[off:off + int(np.prod(ds))].view(ds)

    def gview(self, name):
        return self.view(name, self.grads)

    def pview(self, name):
        """the prepared form of a GEMM weight: BN-folded (towers) / tf32-rounded.  With unrounded operands (h3 / x3) the
        prepared form of a weight WITHOUT batch norm is the weight itself: no copy is kept of those"""
        if self.rounding == 0 and not name.startswith("FirstStage"):
            return self.view(name)
        return self.view(name, self.prep)

    def _arena_off(self, t):
        """float offset of a weight view inside its arena (params and prep share the layout)"""
        for base in (self.prep, self.params):
            d = t.data_ptr() - base.data_ptr()
            if 0 <= d < base.numel() * 4:
                return d // 4
        raise ValueError("not a view of the parameter arenas")

    def _to_dev_layout(self, n, a):
        s, k = self.ptable[n]
        a = np.asarray(a, np.float32)
        assert tuple(a.shape) == tuple(s), (n, a.shape, s)
        if k == "weights":
            if a.ndim == 4:
                a = a.transpose(3, 0, 1, 2).reshape(a.shape[3], -1)       # HWIO -> O,(H,W,I)
            else:
                a = a.T                                                     # [in,out] -> [out,in]
                ds = self.layout[n][2]
                if a.shape[1] != ds[1]:
                    a = np.concatenate([a, np.zeros((a.shape[0], ds[1] - a.shape[1]), np.float32)], 1)
        return torch.from_numpy(np.ascontiguousarray(a))

    def load_params(self, P, slots=None):
        """P: name -> numpy array in TF layout (conv HWIO, fc [in,out]).  The optimizer state is re-initialised (Adam
        moments 0, EMA shadows = the variables, as a fresh TF graph does) except for the variables in `slots`
        (name -> (adam_m, adam_v, ema) in TF layout): a training resume restores those as tf.train.Saver does."""
        for n in self.ptable:
            self.view(n).copy_(self._to_dev_layout(n, P[n]))
        self.ema.copy_(self.params)
        self.adam_m.zero_()
        self.adam_v.zero_()
        for n, (m_, v_, e_) in (slots or {}).items():
            if self.layout[n][0] == "T":
                self.view(n, self.adam_m).copy_(self._to_dev_layout(n, m_))
                self.view(n, self.adam_v).copy_(self._to_dev_layout(n, v_))
                self.view(n, self.ema).copy_(self._to_dev_layout(n, e_))
        self._prepared = False

    def export_params(self, arena=None):
        out = {}
        for n, (s, k) in self.ptable.items():
            t = self.view(n, arena if (arena is not None and self.layout[n][0] == "T") else None).detach().cpu().numpy()
            if k == "weights":
                if len(s) == 4:
                    kk, _, cin, cout = s
                    t = t.reshape(cout, kk, kk, cin).transpose(1, 2, 3, 0)
                else:
                    t = t[:, :s[0]].T
            out[n] = np.ascontiguousarray(t)
        return out

    def export_grads(self):
        return self.export_params(self.grads)

    def load_checkpoint(self, prefix, kind="monopsr", use_ema=False, resume=False):
        """Restore variables from a TensorFlow checkpoint `prefix` (core/tf_checkpoint.py): kind 'monopsr' = a MonoPSR
        training checkpoint (use_ema: the MovingAverageOptimizer shadows, as evaluation does), 'detection' = the
        pre-trained object-detection-API ResNet-101 for both encoders (core/checkpoint_utils.py:64-117).  Variables
        absent from the checkpoint (or of another shape) keep their current values; returns the mapping report."""
        from . import tf_checkpoint
        table = [(n, s, k) for n, (s, k) in self.ptable.items()]
        loaded, report = tf_checkpoint.load_checkpoint(prefix, table, kind=kind, use_ema=use_ema,
                                                        with_slots=resume and kind == "monopsr")
        P = self.export_params()
        P.update(loaded)
        # resume=True (the trainer): Adam moments and EMA shadows come back too -- restarting them at 0 / at the
        # variables would make the first updates after a resume ~3x too large (bias correction uses t = step + 1)
        self.load_params(P, slots=report.get("slots") if resume else None)
        return report

    def save_checkpoint(self, prefix, global_step=None):
        """Write the variables (TF names and layouts) and their EMA shadows as a TensorFlow tensor bundle."""
        from . import tf_checkpoint
        T = self.export_params()
        trainable = set(self.trainable_names)
        for arena, sfx in ((self.ema, tf_checkpoint.EMA_SUFFIX), (self.adam_m, tf_checkpoint.ADAM_M_SUFFIX),
                           (self.adam_v, tf_checkpoint.ADAM_V_SUFFIX)):
            for n, v in self.export_params(arena).items():
                if n in trainable:
                    T[n + sfx] = v
        if global_step is not None:
            T["global_step"] = np.array(int(global_step), np.int32)       # the reference's global_step is int32
        tf_checkpoint.write_bundle(prefix, T, with_data_crc=True)

    # ------------------------------------------------------------------ activations
    def _alloc_activations(self):
        N = self.N
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.dev)
        self.towers = {}
        for enc, nimg, Hin, Win in ((ms.ENCODERS[0], N, ms.CROP, ms.CROP), (ms.ENCODERS[1], 1, ms.FULL_H, ms.FULL_W)):
            H2, W2 = Hin // 2, Win // 2
            h, w = H2 // 2, W2 // 2
            M = nimg * h * w
            T = dict(enc=enc, nimg=nimg, Hin=Hin, Win=Win, H2=H2, W2=W2, h=h, w=w, M=M)
            T["stem"] = e(nimg * H2 * W2, 64)
            T["pool"] = e(M, 64)
            T["g_stem"] = e(nimg * H2 * W2, 64)
            T["g_pool"] = e(M, 64)
            T["tapmask"] = {}
            for rate in (1, 2, 4):
                tm = torch.empty(M, dtype=torch.int16, device=self.dev)
                self._chk(self.L.mpb_build_tapmask(nimg, h, w, 3, 3, rate, _ptr(tm), self._st()), "tapmask")
                T["tapmask"][rate] = tm
            units = []
            cin = 64
            for name, base, nunits, rate in ms.BLOCKS:
                for u in range(1, nunits + 1):
                    scope = "%s/resnet_v1_101/%s/unit_%d/bottleneck_v1" % (enc, name, u)
                    U = dict(scope=scope, cin=cin, base=base, cout=base * 4, rate=rate, proj=(cin != base * 4))
                    U["y1"], U["y2"] = e(M, base), e(M, base)
                    U["g1"], U["g2"] = e(M, base), e(M, base)
                    U["sc"] = e(M, base * 4) if U["proj"] else None
                    U["t"] = e(M, cin) if U["proj"] else None
                    last = (name == "block3" and u == nunits)
                    if last and nimg > 1:
                        U["out"] = None          # written into the concat buffer (crop tower)
                    else:
                        U["out"] = e(M, base * 4)
                    U["g_out"] = e(M, base * 4)
                    # fp32 residual stream: `out` stays unrounded (residual add, ReLU mask), `out_r` is the
                    # tf32-rounded copy that feeds the next unit's GEMMs (halves the forward error)
                    U["out_r"] = None if last else e(M, base * 4)
                    units.append(U)
                    cin = base * 4
            T["units"] = units
            self.towers[enc] = T
        Mc = self.towers[ms.ENCODERS[0]]["M"]          # N*12*12
        self.Mc = Mc
        self.concat = e(Mc, 2048)
        self.towers[ms.ENCODERS[0]]["units"][-1]["out_view"] = (self.concat, 2048)
        self.squashed = e(Mc, 512)
        self.pooled = e(N * 36, 512)
        self.r1 = e(N * 576, 512)
        self.dec = []
        for (blk, cin, cout, M) in (("conv2", 512, 256, N * 576), ("conv3", 256, 128, N * 2304)):
            for i in (1, 2):
                sc = "map_decoder/%s/%s_%d" % (blk, blk, i)
                D = dict(scope=sc, cin=cin if i == 1 else cout, cout=cout, M=M, side=24 if blk == "conv2" else 48)
                D["z"], D["y"] = e(M, cout), e(M, cout)
                D["mean"], D["var"] = e(cout), e(cout)
                D["dz"], D["dy"] = e(M, cout), e(M, cout)
                self.dec.append(D)
        self.r2 = e(N * 2304, 256)
        self.d_r2 = e(N * 2304, 256)
        self.d_r1 = e(N * 576, 512)
        self.bn_scratch = torch.zeros(2 * 512, dtype=torch.float64, device=self.dev)
        self.tm24 = torch.empty(N * 576, dtype=torch.int16, device=self.dev)
        self.tm48 = torch.empty(N * 2304, dtype=torch.int16, device=self.dev)
        self._chk(self.L.mpb_build_tapmask(N, 24, 24, 3, 3, 1, _ptr(self.tm24), self._st()), "tapmask")
        self._chk(self.L.mpb_build_tapmask(N, 48, 48, 3, 3, 1, _ptr(self.tm48), self._st()), "tapmask")
        self.xyz = e(N * 2304, 3)
        self.d_xyz = e(N * 2304, 3)
        # FC stacks
        self.fc = {}
        for key in ("proposal", "regression"):
            F = dict(acc=e(N, 1024), feat=torch.zeros(N, KPAD, dtype=torch.float32, device=self.dev),
                     h0=e(N, 1024), h1=e(N, 1024), d_h1=e(N, 1024), g_h1=e(N, 1024), d_h0=e(N, 1024),
                     g_h0=e(N, 1024), d_feat=e(N, KPAD), g_img=e(N, 1024))
            self.fc[key] = F
        self.d_flat = e(N, 18432)
        self.d_squashed = e(Mc, 512)
        self.g_squashed = e(Mc, 512)
        self.g_fullcrop = e(Mc, 1024)
        self.d_fullfeat = e(self.towers[ms.ENCODERS[1]]["M"], 1024)
        # heads
        self.h = {k: e(*s) for k, s in dict(
            lwh_offs=(N, 3), alpha=(N, 24), cen_y_offs=(N,), cen_z_offs=(N,), lwh=(N, 3), prop_cen_z=(N,),
            prop_cen_y=(N,), cen_x=(N,), cen_y=(N,), cen_z=(N,), centroids=(N, 3), proj_err_norm=(N,),
            depth_global=(N, 2304), losses=(9,), d_lwh_offs=(N, 3), d_alpha=(N, 24), d_cen_y_offs=(N,),
            d_cen_z_offs=(N,), d_prop_y=(N,), d_prop_z=(N,), maskstats=(N + 1,)).items()}
        self.inputs = {}
        torch.cuda.synchronize(self.dev)

    # ------------------------------------------------------------------ helpers
    def _st(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def _chk(self, status, what):
        if status != 0:
            _lib.check(status, what)

    def _cur(self):
        return torch.cuda.current_stream(self.dev)

    @staticmethod
    def set_library_rounding(dev, on):
        """the library's process-wide operand-rounding switch (a __constant__), tracked per device"""
        dev = torch.device(dev)
        key = dev.index or 0
        if Engine._rounding_now.get(key, 1) != int(on):
            with torch.cuda.device(dev):
                torch.cuda.synchronize(dev)
                _lib.check(_lib.load().mpb_set_operand_rounding(int(on)), "mpb_set_operand_rounding")
            Engine._rounding_now[key] = int(on)

    def _enter(self):
        """make the library's operand-rounding switch match this engine (no-op in the common single-engine case)"""
        Engine.set_library_rounding(self.dev, self.rounding)

    def _side(self, stream):
        """context: run on `stream` after everything issued so far on the current stream"""
        if not self.overlap:
            return torch.cuda.stream(self._cur())
        stream.wait_stream(self._cur())
        return torch.cuda.stream(stream)

    def _join(self, *streams):
        if self.overlap:
            for st_ in streams:
                self._cur().wait_stream(st_)

    def s16(self, t):
        """the fp16 hi/lo split copy that shadows fp32 tensor `t` (same storage geometry; views map to views)"""
        st = t.untyped_storage()
        key = st.data_ptr()
        sh = self._s16.get(key)
        if sh is None:
            sh = torch.empty(st.nbytes() // 4, dtype=torch.float32, device=self.dev)
            self._s16[key] = sh
        return torch.as_strided(sh, t.shape, t.stride(), t.storage_offset())

    def _split(self, t, rows, C, ld):
        """h3: write the split copy of a GEMM operand produced by a non-GEMM kernel (pools, resizes, batch norm, FC glue)"""
        if self.h3:
            self._chk(self.L.mpb_split16(rows, C, _ptr(t), ld, _ptr(self.s16(t)), ld, 0, _ptr(self.overflow), self._st()),
                      "split16")

    def check_overflow(self):
        """h3 only: raise if a forward activation left fp16's range since the last check (then use precision='x3')"""
        if self.h3 and int(self.overflow.item()) != 0:
            self.overflow.zero_()
            raise _lib.MpbError("h3 forward: an activation beyond +-65504 was split into fp16 halves; "
                                "results are invalid -- run this model with precision='x3'")

    def _plan_tiles(self, mtiles, ncols, nkb, csk=None):
        """(tile width, split-K) of a FWD / DGRAD launch; rules read off tools/gemm_sweep.py on B200
        (profiles/r1_notes.md).  Wide tiles halve the operand traffic per FLOP and make the main loop MMA-bound,
        but a 128 x 256 tile grid of these layers covers only 36-48 SMs: long reductions are therefore split in
        two over a 2-CTA cluster (a TPC) that reduces through DSMEM; clusters of 3-4 fragment the GPCs."""
        tiles = {bn: mtiles * (ncols // bn) for bn in (256, 128, 64) if ncols % bn == 0}
        lo = int(self.fill * 100)                     # CTAs below which a launch is considered too small
        if (self.csk if csk is None else csk) and nkb >= 16:
            for bn in (256, 128):
                if bn in tiles and bn <= self.csk_bn and lo <= 2 * tiles[bn] <= self.sms * self.ctas_per_sm[bn]:
                    return bn, 2
        for bn in (256, 128):
            if bn in tiles and tiles[bn] >= lo:
                # short reductions are epilogue-bound: once the grid needs a second wave anyway, 64-wide
                # tiles (two CTAs per SM, one's epilogue under the other's main loop) win
                if nkb <= 8 and tiles[bn] > self.sms and self.shortk_bn in tiles:
                    return self.shortk_bn, 1
                return bn, 1
        for bn in (64, 128, 256):
            if bn in tiles:
                return bn, 1
        raise ValueError("no tile width divides %d" % ncols)

    def gemm(self, op, M, H, W, k, dil, Cin, Cout, X, ldx, Wt, ldw, out, ldo, Y=None, ldy=0, tapmask=None,
             shift=None, res=None, ldr=0, mask=None, ldm=0, rowscale=None, colsum=None, relu=0, round_tf32=0,
             atomic=0, ksplit=1, bn=None, out_r=None, ldor=0, out16=False):
        """out16 (h3 only): also write the split copy of `out` -- set it where the result feeds a forward GEMM"""
        p = TcGemmParams()
        p.op, p.H, p.W, p.kh, p.kw, p.dil, p.M, p.Cin, p.Cout = op, H, W, k, k, dil, M, Cin, Cout
        p.X, p.ldx, p.Y, p.ldy, p.Wt, p.ldw, p.out, p.ldo = _ptr(X), ldx, _ptr(Y), ldy, _ptr(Wt), ldw, _ptr(out), ldo
        p.tapmask = _ptr(tapmask)
        p.shift, p.res, p.ldr, p.mask, p.ldm = _ptr(shift), _ptr(res), ldr, _ptr(mask), ldm
        p.rowscale, p.colsum = _ptr(rowscale), _ptr(colsum)
        p.relu, p.round_tf32, p.atomic, p.ksplit = relu, round_tf32, atomic, ksplit
        p.out_r, p.ldor = _ptr(out_r), ldor
        mt = (M + 127) // 128
        if bn is None:
            kdepth = k * k * (Cin if op == TC_FWD else Cout)
            if op == TC_WGRAD:
                bn = 128 if Cin % 128 == 0 else 64
            else:
                bn, ks = self._plan_tiles(mt, Cout if op == TC_FWD else Cin, kdepth // 32,
                                          csk=self.csk_fwd if op == TC_FWD else self.csk)
                if ksplit == 1 and not atomic and M % (H * W) == 0:
                    p.ksplit = ks
        if self.h3 and op == TC_FWD:
            # operands = split copies; results stay unrounded (pools / resizes / batch norm read them), the rounded
            # second copy of the residual stream is replaced by the split copy
            p.round_tf32 = 0
            off = self._arena_off(Wt)
            p.X16 = _ptr(self.s16(X))
            p.W16 = ctypes.c_void_p(self.prep16.data_ptr() + 4 * off)
            p.scale = _ptr(self.w16_inv[off])
            p.overflow = _ptr(self.overflow)
            if out16 or out_r is not None:
                p.out16, p.ldo16 = _ptr(self.s16(out)), ldo
            p.out_r, p.ldor = None, 0
            if not atomic:
                p.ksplit = 1
            if getattr(self, "_record", None) is not None:
                self._record.append((p, bn, 2.0 * M * Cin * Cout * k * k, "h3"))
            self._chk(self.L.mpb_tc_gemm_h3(ctypes.byref(p), bn, self._st()), "mpb_tc_gemm_h3")
            return
        if self.x3 and op == TC_FWD:
            # operands stay unrounded: no rounding in the epilogue, no rounded second copy (callers read `out`)
            p.round_tf32, p.out_r, p.ldor = 0, None, 0
            if not atomic:
                p.ksplit = 1
            bn3 = 128 if Cout % 128 == 0 else 64
            if getattr(self, "_record", None) is not None:
                self._record.append((p, bn3, 2.0 * M * Cin * Cout * k * k, "x3"))
            self._chk(self.L.mpb_tc_gemm_x3(ctypes.byref(p), bn3, self._st()), "mpb_tc_gemm_x3")
            return
        if getattr(self, "_record", None) is not None:
            self._record.append((p, bn, 2.0 * M * Cin * Cout * k * k, "tf32"))
        self._chk(self.L.mpb_tc_gemm(ctypes.byref(p), bn, self._st()), "mpb_tc_gemm")

    def gemm_only_roofline(self, flush=None, iters=10):
        """Time ONLY the tcgen05 GEMM launches of one training step (same arguments, replayed as a CUDA
        graph) and return their algorithmic FLOP rate (2*M*N*K, full-tap convention of SURVEY.md 8d)."""
        self._record = []
        self.forward(train=True)
        self.backward()
        rec, self._record = self._record, None
        torch.cuda.synchronize(self.dev)
        s = torch.cuda.Stream(device=self.dev)
        with torch.cuda.stream(s):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
                fns = {"tf32": self.L.mpb_tc_gemm, "x3": self.L.mpb_tc_gemm_x3, "h3": self.L.mpb_tc_gemm_h3}
                for p, bn, _, kind in rec:
                    self._chk(fns[kind](ctypes.byref(p), bn, self._st()), "mpb_tc_gemm (%s)" % kind)
        torch.cuda.synchronize(self.dev)
        g.replay()
        torch.cuda.synchronize(self.dev)
        ts = []
        for _ in range(iters):
            if flush is not None:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize(self.dev)
            ts.append(a.elapsed_time(b))
        ms_ = float(np.median(ts))
        flop = float(sum(r[2] for r in rec))
        kinds = {}
        for r in rec:
            kinds[r[3]] = kinds.get(r[3], 0) + 1
        return {"ms": ms_, "gflop": flop / 1e9, "tflops": flop / (ms_ * 1e-3) / 1e12, "launches": len(rec), "kinds": kinds}

    def wgrad(self, M, H, W, k, dil, Cin, Cout, X, ldx, dY, ldy, dW, tapmask=None, rowscale=None):
        # widest tile that divides Cin, then enough K slices (RED.ADD into the zeroed gradient arena) to put
        # about one CTA on every SM: 128 x 256 tiles at ~144 CTAs ran 1.7x faster than 64-wide ones
        bn = next(b for b in (256, 128, 64) if Cin % b == 0 and b <= self.wgrad_bn)
        tiles = ((Cout + 127) // 128) * (k * k * Cin // bn)
        nkb = (M + 31) // 32
        ksplit = max(1, min(int(self.wgrad_fill * self.sms + tiles // 2) // tiles, max(1, nkb // 4)))
        self.gemm(TC_WGRAD, M, H, W, k, dil, Cin, Cout, X, ldx, None, k * k * Cin, dW, 0, Y=dY, ldy=ldy,
                  tapmask=tapmask, rowscale=rowscale, atomic=1, ksplit=ksplit, bn=bn)

    # ------------------------------------------------------------------ weight preparation
    def _build_bn_table(self):
        """device table of every frozen-BN conv of both towers (one launch folds / differentiates them all)"""
        self.bnfold, self.bn_row0 = {}, {}
        layers = []
        for enc in ms.ENCODERS:
            for scope, k, cin, cout, _ in ms.conv_layers(enc):
                self.bnfold[scope] = (torch.empty(cout, device=self.dev), torch.empty(cout, device=self.dev))
                layers.append((scope, k * k * cin, cout))
        arr = (BnLayer * len(layers))()
        row2layer = []
        row = 0
        for i, (scope, K, cout) in enumerate(layers):
            b = scope + "/BatchNorm/"
            e = arr[i]
            e.w, e.gamma, e.beta = self.view(scope + "/weights").data_ptr(), self.view(b + "gamma").data_ptr(), self.view(b + "beta").data_ptr()
            e.mean, e.var = self.view(b + "moving_mean").data_ptr(), self.view(b + "moving_variance").data_ptr()
            e.wf, e.scale, e.shift = self.pview(scope + "/weights").data_ptr(), self.bnfold[scope][0].data_ptr(), self.bnfold[scope][1].data_ptr()
            e.dw, e.dbeta, e.dgamma = self.gview(scope + "/weights").data_ptr(), self.gview(b + "beta").data_ptr(), self.gview(b + "gamma").data_ptr()
            e.cout, e.K, e.row0 = cout, K, row
            self.bn_row0[scope] = row
            row2layer += [i] * cout
            row += cout
        self.bn_rows = row
        self.bn_layers = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.dev)
        self.bn_row2layer = torch.tensor(row2layer, dtype=torch.int32, device=self.dev)
        # the two 7x7x3 stems alone (h3: every other tower conv is folded by the split-copy pass)
        stems = [i for i, (scope, K, cout) in enumerate(layers) if K % 32]
        sarr = (BnLayer * len(stems))()
        s2l, srow = [], 0
        for j, i in enumerate(stems):
            ctypes.memmove(ctypes.addressof(sarr[j]), ctypes.addressof(arr[i]), ctypes.sizeof(BnLayer))
            sarr[j].row0 = srow
            s2l += [j] * sarr[j].cout
            srow += sarr[j].cout
        self.bn_stem = (srow, torch.frombuffer(bytearray(bytes(sarr)), dtype=torch.uint8).to(self.dev),
                        torch.tensor(s2l, dtype=torch.int32, device=self.dev))
        # the non-tower GEMM weights only need tf32 rounding: they are contiguous at the end of the arena
        first = min(self.layout[n][1] for n in self.trainable_names if not n.startswith("FirstStage"))
        self.round_off, self.round_len = first, self.n_train - first

    def _build_w16_table(self):
        """h3: device tables of every forward tensor-core GEMM weight (tower convs with their frozen BN, squash, decoder
        convs, the six FC layers) for the one-launch split / scale pass; `w16_inv[arena offset]` = 1 / row scale"""
        self.w16_inv, self.w16_tabs = {}, {}
        tower, head = [], []
        for enc in ms.ENCODERS:
            for scope, k, cin, cout, _ in ms.conv_layers(enc):
                if k * k * cin % 32:
                    continue                                    # the 7x7x3 stem is a SIMT kernel
                b = scope + "/BatchNorm/"
                tower.append((scope + "/weights", self.view(b + "gamma"), self.view(b + "moving_variance"), scope))
        gemm_heads = ["squash/1x1_conv"] + [D["scope"] for D in self.dec] + \
                     ["output/%s_fc/%s_fc/%s" % (a, a, l) for a in ("proposal", "regression") for l in ("img_fc", "fc0", "fc1")]
        for sc in gemm_heads:
            head.append((sc + "/weights", None, None, None))
        for key, rows_ in (("towers", tower), ("head", head)):
            arr = (W16Layer * len(rows_))()
            row2layer, row = [], 0
            for i, (name, gamma, var, scope) in enumerate(rows_):
                _, off, ds = self.layout[name]
                cout, K = ds
                inv = torch.empty(cout, device=self.dev)
                self.w16_inv[off] = inv
                e = arr[i]
                e.w = self.params.data_ptr() + 4 * off
                e.gamma = gamma.data_ptr() if gamma is not None else None
                e.var = var.data_ptr() if var is not None else None
                e.w16 = self.prep16.data_ptr() + 4 * off
                e.inv_scale = inv.data_ptr()
                if scope is not None:        # frozen-BN conv: this pass also folds (wf, scale, shift) -- one read of w
                    b = scope + "/BatchNorm/"
                    e.wf = self.prep.data_ptr() + 4 * off
                    e.scale, e.shift = self.bnfold[scope][0].data_ptr(), self.bnfold[scope][1].data_ptr()
                    e.beta, e.mean = self.view(b + "beta").data_ptr(), self.view(b + "moving_mean").data_ptr()
                e.cout, e.K, e.row0 = cout, K, row
                row2layer += [i] * cout
                row += cout
            # rows in classes by length: the kernel keeps a row in registers between its two passes
            Ks = np.array([arr[i].K for i in row2layer])
            classes, lo = [], 0
            for cap in (256, 1024, 2304, 1 << 30):
                idx = np.nonzero((Ks > lo) & (Ks <= cap))[0].astype(np.int32)
                if len(idx):
                    classes.append((len(idx), torch.from_numpy(idx).to(self.dev), int(min(cap, Ks[idx].max()))))
                lo = cap
            self.w16_tabs[key] = (classes, torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.dev),
                                  torch.tensor(row2layer, dtype=torch.int32, device=self.dev))

    def prepare_weights(self, part="all"):
        """fold frozen BN into the tower convs, tf32-round every GEMM weight (run after each update)."""
        self._enter()
        L, st = self.L, self._st()
        if getattr(self, "bn_layers", None) is None:
            self._build_bn_table()
        if part in ("all", "towers"):
            if self.h3:      # the split-copy pass below folds every tensor-core conv; only the two stems are left
                rows_, tab, r2l = self.bn_stem
                self._chk(L.mpb_fold_bn_multi(rows_, _ptr(tab), _ptr(r2l), BN_EPS_RESNET, st), "fold_bn_multi (stems)")
            else:
                self._chk(L.mpb_fold_bn_multi(self.bn_rows, _ptr(self.bn_layers), _ptr(self.bn_row2layer), BN_EPS_RESNET, st),
                          "fold_bn_multi")
        if part in ("all", "head") and self.rounding:      # (unrounded operands: pview() hands out the weights themselves)
            self._chk(L.mpb_round_copy(self.round_len, _ptr(self.params[self.round_off:]), _ptr(self.prep[self.round_off:]),
                                       st), "round_copy")
        if self.h3:
            if getattr(self, "w16_tabs", None) is None:
                self._build_w16_table()
            for key in ("towers", "head"):
                if part in ("all", key):
                    classes, tab, r2l = self.w16_tabs[key]
                    for nrows, rowlist, cap in classes:
                        self._chk(L.mpb_split16_weights_rows(nrows, _ptr(rowlist), cap, _ptr(tab), _ptr(r2l), BN_EPS_RESNET, st),
                                  "split16_weights")
        if part in ("all", "towers"):
            self._prepared = True

    # ------------------------------------------------------------------ inputs
    def set_inputs(self, S):
        """S: dict of numpy arrays / torch tensors (model_spec.synthetic_sample keys); host->device copy.
        A sample may carry the RAW training inputs `depth_map` (H,W) and `instance_masks` (N,H,W) instead of the three
        ground-truth maps: they are then synthesised on the GPU (core/targets.py; monopsr_model.py:165-203).  Likewise
        the RAW camera image `rgb_image` (H,W,3) instead of `rgb_crops` + `full_img` (monopsr_model.py:128-133,222-233)
        -- the form datasets/kitti_loader.engine_sample produces."""
        if "rgb_image" in S and "rgb_crops" not in S:
            from . import targets
            S = dict(S)
            im = targets.image_inputs(S.pop("rgb_image"), S["boxes_2d_norm"], self.dev,
                                      image_input_shape=(ms.FULL_H * 2, ms.FULL_W * 2), img_roi_size=ms.CROP,
                                      resized_full_img_shape=(ms.FULL_H, ms.FULL_W), mean_sub_type="kitti")
            S["rgb_crops"], S["full_img"] = im["rgb_crops"], im["full_img"]
        if "depth_map" in S and "gt_inst_xyz_maps_local" not in S:
            from . import targets
            S = dict(S)
            maps = targets.gt_maps_from_depth(S.pop("depth_map"), S.pop("instance_masks"), S["boxes_2d"], S["boxes_3d"],
                                              S["est_view_angs"], S["cam_p"], self.dev, roi=ms.CROP,
                                              centroid_type="middle", rotate_view=True)
            S.update(maps)
        for k, v in S.items():
            t = torch.as_tensor(np.asarray(v)) if not isinstance(v, torch.Tensor) else v
            if t.dtype == torch.float64:
                t = t.float()
            if k in self.inputs and self.inputs[k].shape == t.shape:
                self.inputs[k].copy_(t, non_blocking=True)
            elif k in self.inputs and self._captured():
                # the captured graphs (and the heads' pointer table) hold the ADDRESS of the first buffer: a reallocation
                # would leave them reading stale memory
                raise _lib.MpbError("input %r changed shape from %s to %s after the step was captured into a CUDA graph; "
                                    "create a new Engine for another input geometry" % (k, tuple(self.inputs[k].shape),
                                                                                       tuple(t.shape)))
            else:
                self.inputs[k] = t.to(self.dev).contiguous()
        self._heads_io = None

    def _captured(self):
        return any(getattr(self, g, None) is not None for g in ("_graph", "_g_fb", "_g_dp"))

    def heads_io(self):
        if getattr(self, "_heads_io", None) is not None:
            return self._heads_io
        I, h = self.inputs, self.h
        io = HeadsIO()
        io.nbox = self.N
        for k in ("boxes_2d", "cam_p", "class_indices", "mean_lwh", "prop_cen_z_offset", "est_view_angs", "boxes_3d",
                  "gt_alpha_bins", "gt_alpha_regs", "gt_alpha_valid_bins", "gt_view_angs"):
            if k in I:
                setattr(io, k, I[k].data_ptr())
        for k, src in (("gt_xyz_local", "gt_inst_xyz_maps_local"), ("gt_xyz_global", "gt_inst_xyz_maps_global"),
                       ("valid_mask", "gt_valid_mask_maps")):
            if src in I:
                setattr(io, k, I[src].data_ptr())
        for k in ("lwh_offs", "alpha", "cen_y_offs", "cen_z_offs", "lwh", "prop_cen_z", "prop_cen_y", "cen_x", "cen_y",
                  "cen_z", "centroids", "proj_err_norm", "depth_global", "losses", "d_lwh_offs", "d_alpha",
                  "d_cen_y_offs", "d_cen_z_offs", "d_prop_y", "d_prop_z", "maskstats"):
            setattr(io, k, h[k].data_ptr())
        io.xyz_local, io.d_xyz_local = self.xyz.data_ptr(), self.d_xyz.data_ptr()
        io.xyz_loss_mode = 0 if self.xyz_loss_type == "smooth_l1_nonzero" else 1
        io.xyz_loss_weight = self.xyz_loss_weight
        io.feat1, io.ld1 = self.fc["proposal"]["feat"].data_ptr(), KPAD
        io.feat2, io.ld2 = self.fc["regression"]["feat"].data_ptr(), KPAD
        io.d_feat2, io.ldd2 = self.fc["regression"]["d_feat"].data_ptr(), KPAD
        self._heads_io = io
        return io

    # ------------------------------------------------------------------ forward
    def _tower_fwd(self, T, x_in):
        L, st = self.L, self._st()      # evaluated inside the caller's stream context
        enc = T["enc"]
        s0 = enc + "/resnet_v1_101/conv1"
        self._chk(L.mpb_stem_fwd(T["nimg"], T["Hin"], T["Win"], _ptr(x_in), _ptr(self.pview(s0 + "/weights")),
                                 _ptr(self.bnfold[s0][1]), _ptr(T["stem"]), st), "stem_fwd")
        self._chk(L.mpb_maxpool3s2_fwd(T["nimg"], T["H2"], T["W2"], 64, _ptr(T["stem"]), _ptr(T["pool"]), st), "pool1")
        x, ldx = T["pool"], 64
        self._split(T["pool"], T["M"], 64, 64)
        xr = x                      # tf32-rounded view of x (the pooled stem output is already rounded)
        M, h, w = T["M"], T["h"], T["w"]
        for U in T["units"]:
            s, cin, base, cout, rate = U["scope"], U["cin"], U["base"], U["cout"], U["rate"]
            if U["out"] is not None:
                out, ldo = U["out"], cout
            else:
                out, ldo = U["out_view"]
            if U["proj"]:
                self.gemm(TC_FWD, M, h, w, 1, 1, cin, cout, xr, ldx, self.pview(s + "/shortcut/weights"), cin,
                          U["sc"], cout, shift=self.bnfold[s + "/shortcut"][1])
                res, ldr = U["sc"], cout
            else:
                res, ldr = x, ldx
            self.gemm(TC_FWD, M, h, w, 1, 1, cin, base, xr, ldx, self.pview(s + "/conv1/weights"), cin, U["y1"], base,
                      shift=self.bnfold[s + "/conv1"][1], relu=1, round_tf32=1, out16=True)
            self.gemm(TC_FWD, M, h, w, 3, rate, base, base, U["y1"], base, self.pview(s + "/conv2/weights"), 9 * base,
                      U["y2"], base, tapmask=T["tapmask"][rate], shift=self.bnfold[s + "/conv2"][1], relu=1, round_tf32=1,
                      out16=True)
            last_crop = U["out"] is None       # written straight into the concat buffer: rounded (squash operand)
            self.gemm(TC_FWD, M, h, w, 1, 1, base, cout, U["y2"], base, self.pview(s + "/conv3/weights"), base, out, ldo,
                      shift=self.bnfold[s + "/conv3"][1], res=res, ldr=ldr, relu=1, round_tf32=1 if last_crop else 0,
                      out_r=U["out_r"], ldor=cout, out16=last_crop)
            U["x"], U["ldx"], U["xr"], U["o"], U["ldo"] = x, ldx, xr, out, ldo
            x, ldx = out, ldo
            xr = U["out_r"] if (U["out_r"] is not None and not (self.x3 or self.h3)) else out
        return x, ldx

    def _fc_layer(self, x, ldx, K, wname, out, ldo, acc):
        """relu(x @ W^T + b): split-K tcgen05 GEMM into a zeroed accumulator, then bias+ReLU."""
        N = self.N
        acc.zero_()
        nkb = K // 32
        ksplit = max(1, min(8, nkb // 8))
        self.gemm(TC_FWD, N, 1, 1, 1, 1, K, 1024, x, ldx, self.pview(wname + "/weights"), K, acc, 1024, atomic=1,
                  ksplit=ksplit, bn=64)
        self._chk(self.L.mpb_bias_relu(N, 1024, _ptr(acc), 1024, _ptr(self.view(wname + "/biases")), 1, 1, _ptr(out), ldo,
                                       self._st()), "bias_relu")
        self._split(out, N, ldo, ldo)        # (the concat tails of `feat` were written before this layer ran)

    def forward(self, train=True, compute_losses=None, features_only=False):
        """train: the reference's is_training (decoder batch norm with batch statistics + moving-average update, and
        the gradient arena is zeroed for a backward pass); compute_losses (default = train): evaluate the losses in
        the heads kernel -- validation runs an is_training=False graph but still reports losses.
        features_only: stop after the feature extractor (net_builder.extract_features, net_builder.py:17-96) and
        return {'features_for_map': (N,48,48,128), 'features_for_box_3d': (N,6,6,512)}; needs only the inputs
        rgb_crops, full_img and boxes_2d_norm."""
        if compute_losses is None:
            compute_losses = train
        self._enter()
        if not self._prepared:
            self.prepare_weights()
        L, st, N, I = self.L, self._st(), self.N, self.inputs
        Tc, Tf = self.towers[ms.ENCODERS[0]], self.towers[ms.ENCODERS[1]]
        if train and self.overlap:
            with self._side(self.s_wf):      # 400 MB zero-fill of the gradient arena, beside the forward pass
                self._chk(L.mpb_zero_fill(self.n_train, _ptr(self.grads), self._st()), "zero_fill")
            self._grads_zeroed = True
        with self._side(self.s_full):
            ff, _ = self._tower_fwd(Tf, I["full_img"])
            self._chk(L.mpb_crop_pool_fwd(Tf["h"], Tf["w"], 1024, _ptr(ff), N, _ptr(I["boxes_2d_norm"]), 24,
                                          _ptr(self.concat[:, 1024:]), 2048, self._st()), "crop_pool_fwd")
            self._split(self.concat[:, 1024:], self.Mc, 1024, 2048)
        self._tower_fwd(Tc, I["rgb_crops"])
        self._join(self.s_full)
        Mc = self.Mc
        self.gemm(TC_FWD, Mc, 12, 12, 1, 1, 2048, 512, self.concat, 2048, self.pview("squash/1x1_conv/weights"), 2048,
                  self.squashed, 512, shift=self.view("squash/1x1_conv/biases"), relu=1, round_tf32=1)
        self._chk(L.mpb_maxpool2_fwd(N, 12, 12, 512, _ptr(self.squashed), 512, _ptr(self.pooled), 512, st), "pool")
        self._split(self.pooled, N * 36, 512, 512)
        if not features_only:
            with self._side(self.s_fc):
                self._fc_forward()
        s16 = (lambda t: _ptr(self.s16(t))) if self.h3 else (lambda t: None)      # split copies written by the producers
        ovf = _ptr(self.overflow) if self.h3 else None
        self._chk(L.mpb_resize_ac_fwd16(N, 12, 12, 512, _ptr(self.squashed), 24, 24, _ptr(self.r1), s16(self.r1), ovf, st),
                  "resize1")
        x = self.r1
        for i, D in enumerate(self.dec):
            if i == 2:
                self._chk(L.mpb_resize_ac_fwd16(N, 24, 24, 256, _ptr(x), 48, 48, _ptr(self.r2), s16(self.r2), ovf, st),
                          "resize2")
                x = self.r2
            side = D["side"]
            self.gemm(TC_FWD, D["M"], side, side, 3, 1, D["cin"], D["cout"], x, D["cin"], self.pview(D["scope"] + "/weights"),
                      9 * D["cin"], D["z"], D["cout"], tapmask=self.tm24 if side == 24 else self.tm48)
            b = D["scope"] + "/BatchNorm/"
            y16 = s16(D["y"]) if i < 3 else None         # the last decoder output feeds the (SIMT) xyz head only
            if train:      # batch statistics + moving-average update (UPDATE_OPS)
                fwd = L.mpb_bn_train_fwd_fused if self.bn_fused else L.mpb_bn_train_fwd16
                self._chk(fwd(D["M"], D["cout"], _ptr(D["z"]), _ptr(self.view(b + "beta")), BN_EPS_DECODER,
                              _ptr(D["y"]), _ptr(D["mean"]), _ptr(D["var"]),
                              _ptr(self.view(b + "moving_mean")), _ptr(self.view(b + "moving_variance")),
                              BN_DECAY_DECODER, _ptr(self.bn_scratch), y16, ovf, st), "bn_train_fwd")
            else:          # validation / inference graphs are built with is_training=False: moving statistics
                self._chk(L.mpb_bn_infer_fwd16(D["M"], D["cout"], _ptr(D["z"]), _ptr(self.view(b + "beta")),
                                               _ptr(self.view(b + "moving_mean")), _ptr(self.view(b + "moving_variance")),
                                               BN_EPS_DECODER, _ptr(D["y"]), y16, ovf, st), "bn_infer_fwd")
            D["x"] = x
            x = D["y"]
        if features_only:
            return {"features_for_map": x.view(N, 48, 48, 128), "features_for_box_3d": self.pooled.view(N, 6, 6, 512)}
        sx = "output/inst_xyz_map_local/inst_xyz_map_local"
        self._chk(L.mpb_xyzhead_fwd(N, 48, 48, _ptr(x), _ptr(self.view(sx + "/weights")), _ptr(self.view(sx + "/biases")),
                                    _ptr(self.xyz), st), "xyzhead_fwd")
        self._join(self.s_fc)            # the FC stacks ran beside the decoder
        io = self.heads_io()
        self._chk(L.mpb_heads_final(ctypes.byref(io), 1 if compute_losses else 0, st), "heads_final")
        if compute_losses and io.xyz_loss_mode == 1:
            self._pointset_loss()

    def _pointset_loss(self):
        """ChamferDistance / EarthMoversDistance as the training loss of the local xyz map (losses_custom.py:135-198):
        clouds = pred * mask and gt * mask as (N, 2304, 3); loss = weight * sum_b(dist_b) / B / num_boxes
        (loss_builder.py:60-84 then monopsr_model.py:585), B = num_boxes; its gradient goes into d_xyz next to the
        projection / depth terms that mpb_heads_final left there."""
        L, st, N, I = self.L, self._st(), self.N, self.inputs
        n = 2304
        npts = N * n
        if self._ps is None:
            e = lambda *s_, dt=torch.float32: torch.empty(*s_, dtype=dt, device=self.dev)
            ps = dict(p=e(N, n, 3), t=e(N, n, 3), g1=e(N, n, 3), g2=e(N, n, 3))
            if self.xyz_loss_type == "chamfer_dist":
                ps.update(d1=e(N, n), d2=e(N, n), i1=e(N, n, dt=torch.int32), i2=e(N, n, dt=torch.int32), c=e(N, n))
            else:
                ps.update(match=e(N, n, n), cost=e(N))
            self._ps = ps
        ps = self._ps
        scale = self.xyz_loss_weight / float(N) / float(N)
        self._chk(L.mpb_pointset_mask(npts, _ptr(self.xyz), _ptr(I["gt_inst_xyz_maps_local"]), _ptr(I["gt_valid_mask_maps"]),
                                      _ptr(ps["p"]), _ptr(ps["t"]), st), "pointset_mask")
        if self.xyz_loss_type == "chamfer_dist":
            self._chk(L.mpb_nn_distance(N, n, _ptr(ps["p"]), n, _ptr(ps["t"]), _ptr(ps["d1"]), _ptr(ps["i1"]), _ptr(ps["d2"]),
                                        _ptr(ps["i2"]), st), "nn_distance")
            self._chk(L.mpb_pointset_loss_add(npts, _ptr(ps["d1"]), npts, _ptr(ps["d2"]), scale, _ptr(self.h["losses"]), 0, 8,
                                              _ptr(ps["c"]), npts, 1.0, st), "pointset_loss_add")
            self._chk(L.mpb_nn_distance_grad(N, n, _ptr(ps["p"]), n, _ptr(ps["t"]), _ptr(ps["c"]), _ptr(ps["i1"]), _ptr(ps["c"]),
                                             _ptr(ps["i2"]), _ptr(ps["g1"]), _ptr(ps["g2"]), st), "nn_distance_grad")
        else:
            self._chk(L.mpb_approxmatch(N, n, n, _ptr(ps["p"]), _ptr(ps["t"]), _ptr(ps["match"]), None, st), "approxmatch")
            self._chk(L.mpb_matchcost(N, n, n, _ptr(ps["p"]), _ptr(ps["t"]), _ptr(ps["match"]), _ptr(ps["cost"]), st), "matchcost")
            self._chk(L.mpb_pointset_loss_add(N, _ptr(ps["cost"]), 0, None, scale, _ptr(self.h["losses"]), 0, 8, None, 0, 0.0,
                                              st), "pointset_loss_add")
            self._chk(L.mpb_matchcostgrad(N, n, n, _ptr(ps["p"]), _ptr(ps["t"]), _ptr(ps["match"]), _ptr(ps["g1"]),
                                          _ptr(ps["g2"]), st), "matchcostgrad")
        self._chk(L.mpb_pointset_grad_add(npts, _ptr(ps["g1"]), _ptr(I["gt_valid_mask_maps"]), scale, _ptr(self.d_xyz), st),
                  "pointset_grad_add")

    def _fc_forward(self):
        """heads_static, proposal stack, lwh/alpha heads, heads_mid, regression stack, cen_y/cen_z heads"""
        L, st, N = self.L, self._st(), self.N
        io = self.heads_io()
        self._chk(L.mpb_heads_static(ctypes.byref(io), st), "heads_static")
        P, R = self.fc["proposal"], self.fc["regression"]
        p = "output/proposal_fc/proposal_fc"
        self._fc_layer(self.pooled, 18432, 18432, p + "/img_fc", P["feat"], KPAD, P["acc"])
        self._fc_layer(P["feat"], KPAD, KPAD, p + "/fc0", P["h0"], 1024, P["acc"])
        self._fc_layer(P["h0"], 1024, 1024, p + "/fc1", P["h1"], 1024, P["acc"])
        h = self.h
        for nm, n, out in (("output/lwh/lwh", 3, h["lwh_offs"]), ("output/alpha", 24, h["alpha"])):
            self._chk(L.mpb_fc_small_fwd(N, 1024, n, _ptr(P["h1"]), 1024, _ptr(self.view(nm + "/weights")),
                                         _ptr(self.view(nm + "/biases")), _ptr(out), n, st), "fc_small_fwd")
        self._chk(L.mpb_heads_mid(ctypes.byref(io), st), "heads_mid")
        r = "output/regression_fc/regression_fc"
        self._fc_layer(self.pooled, 18432, 18432, r + "/img_fc", R["feat"], KPAD, R["acc"])
        self._fc_layer(R["feat"], KPAD, KPAD, r + "/fc0", R["h0"], 1024, R["acc"])
        self._fc_layer(R["h0"], 1024, 1024, r + "/fc1", R["h1"], 1024, R["acc"])
        for nm, out in (("output/cen_y/cen_y", h["cen_y_offs"]), ("output/cen_z_offs/cen_z", h["cen_z_offs"])):
            self._chk(L.mpb_fc_small_fwd(N, 1024, 1, _ptr(R["h1"]), 1024, _ptr(self.view(nm + "/weights")),
                                         _ptr(self.view(nm + "/biases")), _ptr(out), 1, st), "fc_small_fwd")

    def outputs(self):
        """output_dict (core/constants.py KEY_*) as device tensors."""
        N, h = self.N, self.h
        o = {
            "inst_xyz_map_local": self.xyz.view(N, 48, 48, 3),
            "valid_mask_maps": self.inputs.get("gt_valid_mask_maps"),
            "lwh": h["lwh"], "lwh_offs": h["lwh_offs"], "alpha_bins": h["alpha"][:, :12], "alpha_regs": h["alpha"][:, 12:],
            "view_ang": self.inputs["est_view_angs"].view(N, 1), "prop_cen_z": h["prop_cen_z"].view(N, 1),
            "cen_x": h["cen_x"].view(N, 1), "cen_y": h["cen_y"].view(N, 1), "cen_y_offs": h["cen_y_offs"].view(N, 1),
            "cen_z": h["cen_z"].view(N, 1), "cen_z_offs": h["cen_z_offs"].view(N, 1), "centroids": h["centroids"],
            "proj_err_norm": h["proj_err_norm"], "inst_depth_map_global": h["depth_global"].view(N, 48, 48, 1),
        }
        return o

    def losses(self):
        names = ["inst_xyz_map_local", "lwh_offs", "alpha_bins", "alpha_regs", "cen_z_offs", "cen_y_offs", "proj_err",
                 "inst_depth_map_global", "total_loss"]
        v = self.h["losses"].detach().cpu().numpy()
        return {n: float(x) for n, x in zip(names, v)}

    # ------------------------------------------------------------------ backward
    def _fc_bwd(self, y, ldy, dy, lddy, g, x, ldx, K, wname, dx, lddx, res=None, ldr=0, want_dx=True):
        """backward of relu(x W^T + b): g = relu'(y)*dy; db = colsum(g); dW += g^T x; dx = g W (+res)."""
        N = self.N
        self._chk(self.L.mpb_relu_bwd_colsum(N, 1024, _ptr(y), ldy, _ptr(dy), lddy, _ptr(g), 1024,
                                             _ptr(self.gview(wname + "/biases")), self._st()), "relu_bwd")
        self.wgrad(N, 1, 1, 1, 1, K, 1024, x, ldx, g, 1024, self.gview(wname + "/weights"))
        if want_dx:
            self.gemm(TC_DGRAD, N, 1, 1, 1, 1, K, 1024, g, 1024, self.pview(wname + "/weights"), K, dx, lddx,
                      res=res, ldr=ldr, bn=64 if K % 256 else 256)

    def _fc_backward(self):
        """regression stack -> heads_bwd_mid -> proposal stack; leaves d(pooled) in self.d_flat"""
        L, st, N, h = self.L, self._st(), self.N, self.h
        io = self.heads_io()
        P, R = self.fc["proposal"], self.fc["regression"]
        # ---- regression stack
        r = "output/regression_fc/regression_fc"
        first = True
        for nm, dy in (("output/cen_y/cen_y", h["d_cen_y_offs"]), ("output/cen_z_offs/cen_z", h["d_cen_z_offs"])):
            self._chk(L.mpb_fc_small_bwd(N, 1024, 1, _ptr(R["h1"]), 1024, _ptr(self.view(nm + "/weights")), _ptr(dy), 1,
                                         _ptr(R["d_h1"]), 1024, 0 if first else 1, _ptr(self.gview(nm + "/weights")),
                                         _ptr(self.gview(nm + "/biases")), st), "fc_small_bwd")
            first = False
        self._fc_bwd(R["h1"], 1024, R["d_h1"], 1024, R["g_h1"], R["h0"], 1024, 1024, r + "/fc1", R["d_h0"], 1024)
        self._fc_bwd(R["h0"], 1024, R["d_h0"], 1024, R["g_h0"], R["feat"], KPAD, KPAD, r + "/fc0", R["d_feat"], KPAD)
        self._chk(L.mpb_heads_bwd_mid(ctypes.byref(io), st), "heads_bwd_mid")
        self._fc_bwd(R["feat"], KPAD, R["d_feat"], KPAD, R["g_img"], self.pooled, 18432, 18432, r + "/img_fc",
                     self.d_flat, 18432)
        # ---- proposal stack
        p = "output/proposal_fc/proposal_fc"
        first = True
        for nm, n, dy in (("output/lwh/lwh", 3, h["d_lwh_offs"]), ("output/alpha", 24, h["d_alpha"])):
            self._chk(L.mpb_fc_small_bwd(N, 1024, n, _ptr(P["h1"]), 1024, _ptr(self.view(nm + "/weights")), _ptr(dy), n,
                                         _ptr(P["d_h1"]), 1024, 0 if first else 1, _ptr(self.gview(nm + "/weights")),
                                         _ptr(self.gview(nm + "/biases")), st), "fc_small_bwd")
            first = False
        self._fc_bwd(P["h1"], 1024, P["d_h1"], 1024, P["g_h1"], P["h0"], 1024, 1024, p + "/fc1", P["d_h0"], 1024)
        self._fc_bwd(P["h0"], 1024, P["d_h0"], 1024, P["g_h0"], P["feat"], KPAD, KPAD, p + "/fc0", P["d_feat"], KPAD)
        self._fc_bwd(P["feat"], KPAD, P["d_feat"], KPAD, P["g_img"], self.pooled, 18432, 18432, p + "/img_fc",
                     self.d_flat, 18432, res=self.d_flat, ldr=18432)

    def _tower_bwd(self, T, x_in, ws, on_done=None):
        """T['units'][-1]['g_out'] holds g = dL/d(out)*(out>0) of the last unit (and its d(beta3) is set).
        The data-gradient chain runs on the current stream, the weight gradients on `ws`.
        on_done(ui): called when everything of unit ui (and, with ui = -1, of the stem) has been issued."""
        L = self.L
        M, h, w = T["M"], T["h"], T["w"]
        units = T["units"]
        for ui in range(len(units) - 1, -1, -1):
            U = units[ui]
            s, cin, base, cout, rate = U["scope"], U["cin"], U["base"], U["cout"], U["rate"]
            g = U["g_out"]
            x, ldx = U["x"], U["ldx"]
            f = self.bnfold
            # conv3
            with self._side(ws):
                self.wgrad(M, h, w, 1, 1, base, cout, U["y2"], base, g, cout, self.gview(s + "/conv3/weights"),
                           rowscale=f[s + "/conv3"][0])
                if U["proj"]:
                    self.wgrad(M, h, w, 1, 1, cin, cout, U["xr"], ldx, g, cout, self.gview(s + "/shortcut/weights"),
                               rowscale=f[s + "/shortcut"][0])
            self.gemm(TC_DGRAD, M, h, w, 1, 1, base, cout, g, cout, self.pview(s + "/conv3/weights"), base, U["g2"], base,
                      mask=U["y2"], ldm=base, colsum=self.gview(s + "/conv2/BatchNorm/beta"), round_tf32=1)
            # conv2
            tm = T["tapmask"][rate]
            with self._side(ws):
                self.wgrad(M, h, w, 3, rate, base, base, U["y1"], base, U["g2"], base, self.gview(s + "/conv2/weights"),
                           tapmask=tm, rowscale=f[s + "/conv2"][0])
            self.gemm(TC_DGRAD, M, h, w, 3, rate, base, base, U["g2"], base, self.pview(s + "/conv2/weights"), 9 * base,
                      U["g1"], base, tapmask=tm, mask=U["y1"], ldm=base, colsum=self.gview(s + "/conv1/BatchNorm/beta"),
                      round_tf32=1)
            # conv1 (+ shortcut)
            with self._side(ws):
                self.wgrad(M, h, w, 1, 1, cin, base, U["xr"], ldx, U["g1"], base, self.gview(s + "/conv1/weights"),
                           rowscale=f[s + "/conv1"][0])
            if U["proj"]:
                self.gview(s + "/shortcut/BatchNorm/beta").copy_(self.gview(s + "/conv3/BatchNorm/beta"))
                self.gemm(TC_DGRAD, M, h, w, 1, 1, cin, cout, g, cout, self.pview(s + "/shortcut/weights"), cin, U["t"], cin)
                res, ldr = U["t"], cin
            else:
                res, ldr = g, cout
            if ui > 0:
                prev = units[ui - 1]
                dst, colsum = prev["g_out"], self.gview(prev["scope"] + "/conv3/BatchNorm/beta")
            else:
                dst, colsum = T["g_pool"], None
            self.gemm(TC_DGRAD, M, h, w, 1, 1, cin, base, U["g1"], base, self.pview(s + "/conv1/weights"), cin, dst, cin,
                      res=res, ldr=ldr, mask=x, ldm=ldx, colsum=colsum, round_tf32=1)
            if on_done is not None:
                on_done(ui)
        # stem
        st = self._st()
        s0 = T["enc"] + "/resnet_v1_101/conv1"
        self._chk(L.mpb_maxpool3s2_bwd(T["nimg"], T["H2"], T["W2"], 64, _ptr(T["stem"]), _ptr(T["g_pool"]),
                                       _ptr(T["g_stem"]), st), "pool1_bwd")
        Ms = T["nimg"] * T["H2"] * T["W2"]
        self._chk(L.mpb_relu_bwd_colsum(Ms, 64, _ptr(T["stem"]), 64, _ptr(T["g_stem"]), 64, _ptr(T["g_stem"]), 64,
                                        _ptr(self.gview(s0 + "/BatchNorm/beta")), st), "stem_colsum")
        self._chk(L.mpb_stem_wgrad(T["nimg"], T["Hin"], T["Win"], _ptr(x_in), _ptr(T["g_stem"]), _ptr(self.bnfold[s0][0]),
                                   _ptr(self.gview(s0 + "/weights")), st), "stem_wgrad")
        if on_done is not None:
            on_done(-1)

    def backward(self):
        """head part (FC stacks, decoder, squash: every gradient outside the towers), then the two towers"""
        self._backward_head()
        head_opt = None
        if self.early_opt:
            # every gradient outside the towers is final (FC stacks on s_fc, decoder wgrads on s_wf, the rest here): the
            # train-op of those variables may run under the towers' backward pass.  Started at once it competes with the
            # busiest part of that pass for HBM (measured slower); started when the full-image tower's chain reaches
            # unit `early_opt_unit` it fills the thin tail of the pass (block1 / block2 and the stems) instead
            evs = [s_.record_event() for s_ in (self._cur(), self.s_wf, self.s_fc)]

            def launch_head_opt():
                with torch.cuda.stream(self.s_opt):
                    for e in evs:
                        self.s_opt.wait_event(e)
                    self.optimizer_step(1.0, part="head")
                    self.prepare_weights(part="head")
            if self.early_opt_unit is None:
                launch_head_opt()
            else:
                head_opt = (self.early_opt_unit, launch_head_opt)
        self._backward_towers(head_opt=head_opt)

    def _backward_head(self, join=False):
        L, st, N, I, h = self.L, self._st(), self.N, self.inputs, self.h
        if self._grads_zeroed:
            self._join(self.s_wf)            # the zero-fill of the gradient arena ran beside the forward pass
        else:
            self.grads.zero_()
        self._grads_zeroed = False
        io = self.heads_io()
        P, R = self.fc["proposal"], self.fc["regression"]
        with self._side(self.s_fc):
            self._fc_backward()
        # ---- map decoder
        sx = "output/inst_xyz_map_local/inst_xyz_map_local"
        D = self.dec[3]
        with self._side(self.s_wf):      # weight gradient off the decoder's data-gradient chain
            self._chk(L.mpb_xyzhead_wgrad(N, 48, 48, _ptr(D["y"]), _ptr(self.d_xyz), _ptr(self.gview(sx + "/weights")),
                                          _ptr(self.gview(sx + "/biases")), self._st()), "xyzhead_wgrad")
        self._chk(L.mpb_xyzhead_dgrad(N, 48, 48, _ptr(self.view(sx + "/weights")), _ptr(self.d_xyz), _ptr(D["dy"]), st),
                  "xyzhead_dgrad")
        for i in (3, 2, 1, 0):
            D = self.dec[i]
            side, b = D["side"], D["scope"] + "/BatchNorm/"
            tm = self.tm24 if side == 24 else self.tm48
            bwd = L.mpb_bn_train_bwd_fused if self.bn_fused else L.mpb_bn_train_bwd
            self._chk(bwd(D["M"], D["cout"], _ptr(D["z"]), _ptr(D["mean"]), _ptr(D["var"]), BN_EPS_DECODER,
                          _ptr(D["y"]), _ptr(D["dy"]), _ptr(D["dz"]), _ptr(self.gview(b + "beta")),
                          _ptr(self.bn_scratch), st), "bn_train_bwd")
            with self._side(self.s_wf):
                self.wgrad(D["M"], side, side, 3, 1, D["cin"], D["cout"], D["x"], D["cin"], D["dz"], D["cout"],
                           self.gview(D["scope"] + "/weights"), tapmask=tm)
            if i in (3, 1):
                dst = self.dec[i - 1]["dy"]
            elif i == 2:
                dst = self.d_r2
            else:
                dst = self.d_r1
            self.gemm(TC_DGRAD, D["M"], side, side, 3, 1, D["cin"], D["cout"], D["dz"], D["cout"],
                      self.pview(D["scope"] + "/weights"), 9 * D["cin"], dst, D["cin"], tapmask=tm)
            if i == 2:
                self._chk(L.mpb_resize_ac_bwd(N, 24, 24, 256, _ptr(self.d_r2), 48, 48, _ptr(self.dec[1]["dy"]), st), "resize2_bwd")
        self._chk(L.mpb_resize_ac_bwd(N, 12, 12, 512, _ptr(self.d_r1), 24, 24, _ptr(self.d_squashed), st), "resize1_bwd")
        self._join(self.s_fc)            # d(pooled) from the FC stacks
        self._chk(L.mpb_maxpool2_bwd(N, 12, 12, 512, _ptr(self.squashed), 512, _ptr(self.d_flat), 512, _ptr(self.d_squashed),
                                     512, 1, st), "pool_bwd")
        # ---- squash
        Mc = self.Mc
        self._chk(L.mpb_relu_bwd_colsum(Mc, 512, _ptr(self.squashed), 512, _ptr(self.d_squashed), 512, _ptr(self.g_squashed),
                                        512, _ptr(self.gview("squash/1x1_conv/biases")), st), "squash_relu_bwd")
        self.wgrad(Mc, 12, 12, 1, 1, 2048, 512, self.concat, 2048, self.g_squashed, 512, self.gview("squash/1x1_conv/weights"))
        Tc, Tf = self.towers[ms.ENCODERS[0]], self.towers[ms.ENCODERS[1]]
        lastc, lastf = Tc["units"][-1], Tf["units"][-1]
        wsq = self.pview("squash/1x1_conv/weights")
        self.gemm(TC_DGRAD, Mc, 12, 12, 1, 1, 1024, 512, self.g_squashed, 512, wsq, 2048, lastc["g_out"], 1024,
                  mask=self.concat, ldm=2048, colsum=self.gview(lastc["scope"] + "/conv3/BatchNorm/beta"), round_tf32=1)
        self.gemm(TC_DGRAD, Mc, 12, 12, 1, 1, 1024, 512, self.g_squashed, 512, wsq.view(-1)[1024:], 2048,
                  self.g_fullcrop, 1024)
        if join:        # the gradients of the head variables are complete on THIS stream (bucketed all-reduce)
            self._join(self.s_wf, self.s_fc)

    # ---- data parallelism: the tower gradients leave in buckets, under the rest of the backward pass
    def _dp_plan(self, nbuckets):
        """Buckets of the towers' gradient arena in the order the backward pass completes them.  Bucket j = the units
        cut[j] <= u < cut[j-1] of BOTH towers (the same units of the two towers finish at about the same time; the last
        bucket ends with the stems): per tower one contiguous arena range and one contiguous range of rows of the
        frozen-BN table.  Cuts are placed so that every bucket carries about the same number of bytes."""
        if getattr(self, "bn_layers", None) is None:
            self._build_bn_table()
        towers = []
        for ti, enc in enumerate(ms.ENCODERS):
            units = self.towers[enc]["units"]
            first = lambda U: U["scope"] + ("/shortcut" if U["proj"] else "/conv1")
            start = self.layout[enc + "/resnet_v1_101/conv1/weights"][1]
            end = self.layout[ms.ENCODERS[1] + "/resnet_v1_101/conv1/weights"][1] if ti == 0 else self.round_off
            offs = [self.layout[first(U) + "/weights"][1] for U in units] + [end]
            rows = [self.bn_row0[first(U)] for U in units]
            row_start = self.bn_row0[enc + "/resnet_v1_101/conv1"]
            row_end = self.bn_row0[ms.ENCODERS[1] + "/resnet_v1_101/conv1"] if ti == 0 else self.bn_rows
            towers.append(dict(start=start, offs=offs, rows=rows + [row_end], row_start=row_start))
        t0 = towers[0]
        nu = len(t0["offs"]) - 1
        total = t0["offs"][-1] - t0["start"]
        cuts, acc, hi = [], 0, nu
        for u in range(nu - 1, 0, -1):
            acc = t0["offs"][hi] - t0["offs"][u]
            if len(cuts) < nbuckets - 1 and acc >= total / nbuckets:
                cuts.append(u)
                hi = u
        plan = []
        prev = nu
        for c in cuts + [None]:
            rng, rws = [], []
            for t in towers:
                a = t["start"] if c is None else t["offs"][c]
                rng.append((a, t["offs"][prev]))
                rws.append((t["row_start"] if c is None else t["rows"][c], t["rows"][prev]))
            plan.append(dict(unit=-1 if c is None else c, ranges=rng, rows=rws))
            prev = c
        return plan

    def _dp_issue(self, bucket, events):
        """d(gamma) of the bucket's layers, then its all-reduce, on the side stream s_ar -- after the events that mark
        the bucket's gradients final on the data-gradient and weight-gradient streams of both towers"""
        import torch.distributed as dist
        with torch.cuda.stream(self.s_ar):
            for ev in events:
                self.s_ar.wait_event(ev)
            for r0, r1 in bucket["rows"]:
                self._chk(self.L.mpb_bn_param_grad_range(r0, r1, _ptr(self.bn_layers), _ptr(self.bn_row2layer), BN_EPS_RESNET,
                                                         self._st()), "bn_param_grad_range")
            for a, b in bucket["ranges"]:
                self._dp_works.append(dist.all_reduce(self.grads[a:b], op=dist.ReduceOp.SUM, async_op=True))

    def _backward_towers(self, dp_plan=None, head_opt=None):
        L, st, N, I = self.L, self._st(), self.N, self.inputs
        Tc, Tf = self.towers[ms.ENCODERS[0]], self.towers[ms.ENCODERS[1]]
        lastf = Tf["units"][-1]
        if dp_plan is not None:
            by_unit = {bk["unit"]: j for j, bk in enumerate(dp_plan)}
            marks = {}

            def cb(ti, ws):
                def on_done(ui):
                    j = by_unit.get(ui)
                    if j is None:
                        return
                    marks.setdefault(j, []).extend([self._cur().record_event(), ws.record_event()])
                    if ti == 0:                       # the crop tower is issued last: both towers' marks exist now
                        self._dp_issue(dp_plan[j], marks[j])
                return on_done
            with self._side(self.s_full):
                self._chk(L.mpb_crop_pool_bwd(Tf["h"], Tf["w"], 1024, _ptr(lastf["o"]), N, _ptr(I["boxes_2d_norm"]), 24,
                                              _ptr(self.g_fullcrop), 1024, _ptr(self.d_fullfeat), self._st()), "crop_pool_bwd")
                self._chk(L.mpb_relu_bwd_colsum(Tf["M"], 1024, _ptr(lastf["o"]), 1024, _ptr(self.d_fullfeat), 1024,
                                                _ptr(lastf["g_out"]), 1024,
                                                _ptr(self.gview(lastf["scope"] + "/conv3/BatchNorm/beta")), self._st()),
                          "full_relu_bwd")
                self._tower_bwd(Tf, I["full_img"], self.s_wf, on_done=cb(1, self.s_wf))
            self._tower_bwd(Tc, I["rgb_crops"], self.s_wc, on_done=cb(0, self.s_wc))
            self._join(self.s_full, self.s_wf, self.s_wc)
            return
        with self._side(self.s_full):
            self._chk(L.mpb_crop_pool_bwd(Tf["h"], Tf["w"], 1024, _ptr(lastf["o"]), N, _ptr(I["boxes_2d_norm"]), 24,
                                          _ptr(self.g_fullcrop), 1024, _ptr(self.d_fullfeat), self._st()), "crop_pool_bwd")
            self._chk(L.mpb_relu_bwd_colsum(Tf["M"], 1024, _ptr(lastf["o"]), 1024, _ptr(self.d_fullfeat), 1024,
                                            _ptr(lastf["g_out"]), 1024,
                                            _ptr(self.gview(lastf["scope"] + "/conv3/BatchNorm/beta")), self._st()),
                      "full_relu_bwd")
            cb = None
            if head_opt is not None:
                unit, launch = head_opt

                def cb(ui):
                    if ui == unit:
                        self.s_opt.wait_event(self._cur().record_event())
                        launch()
            self._tower_bwd(Tf, I["full_img"], self.s_wf, on_done=cb)
        self._tower_bwd(Tc, I["rgb_crops"], self.s_wc)
        self._join(self.s_full, self.s_wf, self.s_wc)
        # d(gamma) of every frozen BN from (w, dw, dbeta), both towers in one launch
        self._chk(L.mpb_bn_param_grad_multi(self.bn_rows, _ptr(self.bn_layers), _ptr(self.bn_row2layer), BN_EPS_RESNET, st),
                  "bn_param_grad_multi")

    # ------------------------------------------------------------------ train-op
    @classmethod
    def from_config(cls, device, config, **kw):
        """Engine for a parsed yaml config (core/config_utils.parse_yaml_config): the xyz-map loss type / weight and the
        train-op hyper-parameters come from the config instead of the built-in monopsr_model_000 defaults."""
        from . import config_utils
        config_utils.validate_for_engine(config)
        xyz = list(config.model_config.loss_config.inst_xyz_map_local)
        eng = cls(device, xyz_loss=(xyz[0], float(xyz[1])), **kw)
        eng.configure(config.train_config)
        return eng

    def configure(self, train_config):
        """optimizer_builder.build / _create_learning_rate (builders/optimizer_builder.py:23-118) on
        train_config.optimizer.adam_optimizer: exponential_decay (staircase or not) or constant learning rate, the EMA
        decay of MovingAverageOptimizer (use_moving_average: false -> the shadows simply track the variables)."""
        a = train_config.optimizer.adam_optimizer
        kind = getattr(a, "learning_rate_type", "exponential_decay")
        if kind == "exponential_decay":
            self.lr_initial, self.lr_decay_steps = float(a.initial_learning_rate), int(a.decay_steps)
            self.lr_decay_factor, self.lr_staircase = float(a.decay_factor), bool(getattr(a, "staircase", False))
        elif kind == "constant":
            self.lr_initial, self.lr_decay_steps, self.lr_decay_factor, self.lr_staircase = float(a.learning_rate), 1, 1.0, True
        else:
            raise NotImplementedError("learning_rate_type %r" % kind)
        self.ema_decay = float(a.moving_average_decay) if getattr(a, "use_moving_average", False) else 0.0
        self._graph = self._g_fb = self._g_dp = None          # (the EMA decay is baked into a captured train-op launch

    def learning_rate(self, step):
        """tf.train.exponential_decay (yaml:145-150): lr * factor ^ (step / decay_steps), floored when staircase"""
        e = step / float(self.lr_decay_steps)
        return self.lr_initial * (self.lr_decay_factor ** (math.floor(e) if self.lr_staircase else e))

    def set_hyper(self, step):
        t = step + 1
        lr_t = self.learning_rate(step) * math.sqrt(1.0 - ADAM_BETA2 ** t) / (1.0 - ADAM_BETA1 ** t)
        # ring of pinned slots with an event each: the copy of step k may still be queued when the host prepares
        # step k + 1 (graph replays are asynchronous), so one reused buffer could hand step k a later lr_t
        ring = getattr(self, "_hyper_ring", None)
        if ring is None:
            ring = self._hyper_ring = [(torch.zeros(4, pin_memory=True), torch.cuda.Event()) for _ in range(8)]
            self._hyper_i = 0
        host, ev = ring[self._hyper_i % len(ring)]
        self._hyper_i += 1
        ev.synchronize()                     # (no-op unless the host is 8 steps ahead of the device)
        host[0] = lr_t
        self.hyper.copy_(host, non_blocking=True)
        ev.record(self._cur())

    def optimizer_step(self, grad_scale=1.0, part="all"):
        c0, nc, t0, nt = self.opt_parts[part]
        chunk_bytes = ctypes.sizeof(OptChunk)
        self._chk(self.L.mpb_opt_step_range(nc, ctypes.c_void_p(self.opt_chunks.data_ptr() + c0 * chunk_bytes), t0, nt,
                                            _ptr(self.params), _ptr(self.grads), _ptr(self.adam_m), _ptr(self.adam_v),
                                            _ptr(self.ema), ctypes.c_void_p(self.norm2.data_ptr() + 4 * c0), _ptr(self.hyper),
                                            grad_scale, CLIP_GRADIENT_NORM, ADAM_BETA1,
                                            ADAM_BETA2, ADAM_EPSILON, self.ema_decay, self._st()), "opt_step")
        self._prepared = False

    def allreduce_grads(self):
        """the one collective of the path: sum the flat fp32 gradient arena over the NVLink domain"""
        from . import dp
        return int(round(1.0 / dp.allreduce_flat(self.grads)))

    def train_step_eager(self):
        """forward + backward + (all-reduce) + train-op, launched kernel by kernel."""
        self.forward(train=True)
        self.backward()
        world = self.allreduce_grads()
        self.optimizer_step(1.0 / world)
        self.prepare_weights()

    def train_step(self, S=None):
        """One training step on sample S (host arrays are copied in).  The kernel sequence is captured
        into a CUDA graph on first use and replayed afterwards."""
        self._enter()
        if S is not None:
            self.set_inputs(S)
        self.set_hyper(self.step_count)
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if distributed and int(os.environ.get("MPB_DP_GRAPH", "0")):
            # opt-in (MPB_DP_GRAPH=1): ONE graph for the whole data-parallel step, NCCL all-reduces captured inside it:
            # the head bucket leaves under the towers' backward pass, the tower gradients leave in MPB_DP_BUCKETS buckets
            # (d(gamma) + all-reduce on a side stream) as the backward pass completes them.  Measured on 2 B200: bit-
            # identical replicas, 9.90 vs 9.86 ms/step for the default below -- the all-reduce kernels slow the GEMMs
            # they run beside by as much as the overlap hides (profiles/r2_notes.md), so it is not the default
            if getattr(self, "_g_dp", None) is None:
                self._capture_dp()
            self._g_dp.replay()
        elif distributed:
            # four graphs around two NCCL all-reduces issued from the host: the head bucket (arena tail, 45 %) is
            # reduced under the towers' backward pass, the tower bucket under the train-op of the head variables
            # (NVLink traffic beside an HBM-bound pass: different resources), then the towers' train-op
            if getattr(self, "_g_fb", None) is None:
                self._capture_split()
            from . import dp
            self._g_fb.replay()
            head = dp.allreduce_start(self.grads[self.round_off:])     # on NCCL's stream, beside the towers' backward
            self._g_bt.replay()
            dp.allreduce_finish(head)
            if self._g_opt_t0 is not None:
                # tower gradients in two buckets: the first travels beside the head variables' train-op, the second
                # beside the first tower's train-op
                t0 = dp.allreduce_start(self.grads[:self.tower_mid])
                t1 = dp.allreduce_start(self.grads[self.tower_mid:self.round_off])
                self._g_opt_head.replay()
                dp.allreduce_finish(t0)
                self._g_opt_t0.replay()
                dp.allreduce_finish(t1)
                self._g_opt.replay()
            else:
                towers = dp.allreduce_start(self.grads[:self.round_off])   # beside the head variables' train-op
                self._g_opt_head.replay()
                dp.allreduce_finish(towers)
                self._g_opt.replay()
        else:
            if getattr(self, "_graph", None) is None:
                self._capture()
            self._graph.replay()
        self.step_count += 1

    def _warm(self):
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            self.forward(train=True)
            self.backward()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)

    def _capture(self):
        if not self._prepared:
            self.prepare_weights()
        self._warm()
        g = torch.cuda.CUDAGraph()
        c0 = _lib.launch_count()
        # measured: stepping the head variables under the towers' backward pass is ~1.5 % SLOWER than one train-op
        # at the end (the HBM-bound Adam pass slows the concurrent GEMMs more than the overlap saves): off
        early = int(os.environ.get("MPB_EARLY_OPT", "0")) != 0
        u = os.environ.get("MPB_EARLY_OPT_UNIT", "")
        self.early_opt_unit = int(u) if u != "" else None
        with torch.cuda.graph(g, stream=self.s_main, capture_error_mode="thread_local"):
            self.forward(train=True)
            self.early_opt = early
            self.backward()
            self.early_opt = False
            if early:
                self.optimizer_step(1.0, part="towers")
                self.prepare_weights(part="towers")
                self._join(self.s_opt)
            else:
                self.optimizer_step(1.0)
                self.prepare_weights()
        self.launches_per_step = _lib.launch_count() - c0
        self._graph = g

    def dp_step_eager(self, nbuckets=None):
        """the data-parallel step, launch by launch (also what _capture_dp records): forward, head backward, head bucket
        all-reduce || towers' backward with bucketed all-reduces, train-op on the mean gradient"""
        import torch.distributed as dist
        world = dist.get_world_size()
        plan = self._dp_plan(nbuckets or int(os.environ.get("MPB_DP_BUCKETS", "4")))
        self.forward(train=True)
        self._backward_head(join=True)
        self._dp_works = []
        self._dp_issue(dict(rows=[], ranges=[(self.round_off, self.n_train)]), [self._cur().record_event()])
        self._backward_towers(dp_plan=plan)
        for w in self._dp_works:
            w.wait()                                   # the current stream waits for NCCL's
        self._cur().wait_stream(self.s_ar)
        self._dp_works = []
        self.optimizer_step(1.0 / world)
        self.prepare_weights()

    def _capture_dp(self):
        import torch.distributed as dist
        if not self._prepared:
            self.prepare_weights()
        warm = torch.zeros(8, device=self.dev)
        dist.all_reduce(warm)                          # communicator set-up outside the capture
        self._warm()
        g = torch.cuda.CUDAGraph()
        c0 = _lib.launch_count()
        with torch.cuda.graph(g, stream=self.s_main, capture_error_mode="thread_local"):
            self.dp_step_eager()
        self.launches_per_step = _lib.launch_count() - c0
        self._g_dp = g

    def _capture_split(self):
        import torch.distributed as dist
        if not self._prepared:
            self.prepare_weights()
        self._warm()
        # three graphs around two gradient buckets: [forward + head backward] -> all-reduce of the head variables'
        # gradients (arena tail, 45 %) UNDER [towers' backward] -> all-reduce of the tower gradients -> [train-op]
        g1, g2, g3, g4 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        c0 = _lib.launch_count()
        with torch.cuda.graph(g1, stream=self.s_main, capture_error_mode="thread_local"):
            self.forward(train=True)
            self._backward_head(join=True)
        with torch.cuda.graph(g2, stream=self.s_main, capture_error_mode="thread_local"):
            self._backward_towers()
        scale = 1.0 / dist.get_world_size()
        with torch.cuda.graph(g4, stream=self.s_main, capture_error_mode="thread_local"):
            self.optimizer_step(scale, part="head")
            self.prepare_weights(part="head")
        self._g_opt_t0 = None
        if int(os.environ.get("MPB_DP_TOWER_SPLIT", "1")):
            self._g_opt_t0 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._g_opt_t0, stream=self.s_main, capture_error_mode="thread_local"):
                self.optimizer_step(scale, part="tower0")
        with torch.cuda.graph(g3, stream=self.s_main, capture_error_mode="thread_local"):
            self.optimizer_step(scale, part="towers" if self._g_opt_t0 is None else "tower1")
            self.prepare_weights(part="towers")
        self.launches_per_step = _lib.launch_count() - c0
        self._g_fb, self._g_bt, self._g_opt_head, self._g_opt = g1, g2, g4, g3
