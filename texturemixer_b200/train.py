"""One TextureMixer train step on the device: the inner loop of
`run.train_TextureMixer` (run.py:426-514) at lod 0 - permutation sampling, the three
critic updates, the encoder/generator update, the EMA of the inference copies.

    for repeat in range(minibatch_repeats):                      # run.py:510
        run([D_rec_train_op, D_interp_train_op, D_blend_train_op])    # critics, pre-step E/G   (run.py:511)
        run([EG_train_op])                                            # E_zg, E_zl, G, post-step critics (:512)
        run([Es_zg_update_op, Es_zl_update_op, Gs_update_op])          # EMA beta = 0.999         (:513, :233)

Data parallel like the reference: every rank holds all nine networks, takes its contiguous share of
the global batch, and `Optimizer.apply_updates` sums each network's flat gradient across ranks with one NCCL
all-reduce, scales by 1/world and steps Adam(beta1 0, beta2 0.99, eps 1e-8; config.py:84-87).

gram_weight is 0 (VGG-19 weights not redistributable - stated deviation).  Progressive growing
(SURVEY N1): `TrainingSchedule` (run.py:187-226) gives the level of detail per step, `process_reals`
(run.py:68-102) fades / upscales the reals, and `Trainer.step(..., lod=...)` assigns it to every network
(run.py:310) before the losses are evaluated."""
import gc
import os

import numpy as np
import torch

from . import _lib, interp, loss, parallel
from .network import Network
from .optim import GradientBucket, Optimizer
from .runtime import Runtime

NET_FUNCS = dict(E_zg='networks.E_zg', E_zl='networks.E_zl', G='networks.G_res', D_rec='networks.D_patch',
                 D_interp='networks.D_patch', D_blend='networks.D_patch')


def default_config(train_size=128, fmap_base=1024, fmap_max=512, latent_channels=128, scale_h=3, scale_w=3):
    """The values of config.py:39-87 that the hot path reads."""
    latent_res = train_size // 4
    enc = dict(fmap_base=fmap_base, fmap_max=fmap_max, latent_channels=latent_channels, use_pixelnorm=False,
               tanh_at_end=False)
    return dict(
        resolution=train_size, latent_res=latent_res, scale_h=scale_h, scale_w=scale_w,
        E_zg=dict(enc), E_zl=dict(enc, latent_res=latent_res),
        G=dict(fmap_base=fmap_base, fmap_max=fmap_max, latent_res=latent_res, latent_channels=latent_channels,
               use_pixelnorm=False, tanh_at_end=True),
        D_rec=dict(fmap_base=fmap_base, fmap_max=fmap_max, latent_res=-1),
        D_interp=dict(fmap_base=fmap_base, fmap_max=fmap_max, latent_res=-1),
        D_blend=dict(fmap_base=fmap_base, fmap_max=fmap_max, latent_res=-1),
        opt=dict(beta1=0.0, beta2=0.99, epsilon=1e-8), lrate=0.0015, ema_beta=0.999,
        cuda_graphs='step',   # 'step': the whole step replays as CUDA graphs (Trainer._capture_step); 'critics': only
        #                       the critic evaluations (GraphedCritic); False: every launch from Python
        crop_aware=True,      # G_fcn decodes only the latent window each random_crop depends on (loss.crop_window)
        loss=dict(rec_G_weight=1.0, pixel_weight=200.0, interp_G_weight=1.0, blend_interp_G_weight=1.0),
        levels=int(np.log2(latent_res)),
        zg_interp_variational='hard', zl_interp_variational='permutational',      # config.py:61-62 (loss.interp_modes)
        block_size=0, perm=False,                              # config.py:69-70: the permutation sampler's variants
        lr_mirror_augment=False, ud_mirror_augment=False)      # config.py:96 (Trainer.step_from_dataset)


class TrainingSchedule:
    """run.py:187-226: level of detail, resolution, minibatch and learning rate as functions of the images shown."""

    def __init__(self, cur_nimg, resolution_log2, num_gpus=1, lod_initial_resolution=4, lod_training_kimg=1500,
                 lod_transition_kimg=1500, minibatch_base=16, minibatch_dict=None, max_minibatch_per_gpu=None,
                 lrate_base=0.001, lrate_dict=None, tick_kimg_base=1, tick_kimg_dict=None):
        minibatch_dict, max_minibatch_per_gpu = minibatch_dict or {}, max_minibatch_per_gpu or {}
        lrate_dict, tick_kimg_dict = lrate_dict or {}, tick_kimg_dict or {}
        self.kimg = cur_nimg / 1000.0
        phase_dur = lod_training_kimg + lod_transition_kimg
        phase_idx = int(np.floor(self.kimg / phase_dur)) if phase_dur > 0 else 0
        phase_kimg = self.kimg - phase_idx * phase_dur
        self.lod = resolution_log2
        self.lod -= np.floor(np.log2(lod_initial_resolution))
        self.lod -= phase_idx
        if lod_transition_kimg > 0:
            self.lod -= max(phase_kimg - lod_training_kimg, 0.0) / lod_transition_kimg
        self.lod = max(self.lod, 0.0)
        self.resolution = 2 ** (resolution_log2 - int(np.floor(self.lod)))
        self.minibatch = minibatch_dict.get(self.resolution, minibatch_base)
        self.minibatch -= self.minibatch % num_gpus
        if self.resolution in max_minibatch_per_gpu:
            self.minibatch = min(self.minibatch, max_minibatch_per_gpu[self.resolution] * num_gpus)
        self.lrate = lrate_dict.get(self.resolution, lrate_base)
        self.tick_kimg = tick_kimg_dict.get(self.resolution, tick_kimg_base)


def process_reals(x, lod, lr_mirror_augment=False, ud_mirror_augment=False, drange_data=(0, 255), drange_net=(-1, 1),
                  rng=None):
    """run.py:68-102 on the device.  x: [n,C,r,r] uint8 or float device tensor at the dataset's CURRENT resolution
    r = R / 2^floor(lod) (dataset.configure, run.py:430).  -> (reals_fade, reals_orig), fp32 [n,C,R,R]:
    dynamic range (misc.py:38-43: x * scale + bias), optional mirror augmentation (one uniform per sample,
    flipped when >= 0.5), FadeLOD (lerp towards the 2x2 box-filtered image by lod - floor(lod)), UpscaleLOD
    (nearest-neighbour by 2^floor(lod))."""
    import ctypes as C
    from .networks import _lerp_lod, _pool_image, _upscale_image
    rt = Runtime.get(x.device)
    x = x.to(torch.float32).contiguous()
    n, c, h, w = x.shape
    scale = (np.float32(drange_net[1]) - np.float32(drange_net[0])) / (np.float32(drange_data[1]) - np.float32(drange_data[0]))
    bias = np.float32(drange_net[0]) - np.float32(drange_data[0]) * scale
    if tuple(drange_data) != tuple(drange_net):
        y = rt.empty(n, c, h, w)
        _lib.check(rt.lib.tmx_convert_output(rt.handle, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), n * c, h, w,
                                             float(scale), float(bias), 1, 0, rt.stream()), 'tmx_convert_output')
        x = y
    for on, axis in ((lr_mirror_augment, 3), (ud_mirror_augment, 2)):
        if on:
            rng = rng or np.random
            flip = rng.uniform(0.0, 1.0, n) >= 0.5                       # tf.where(mask < 0.5, x, reverse(x))
            length = x.shape[axis]
            idx = np.where(flip[:, None], np.arange(length - 1, -1, -1)[None], np.arange(length)[None]).astype(np.int32)
            idx = torch.from_numpy(np.ascontiguousarray(idx)).to(rt.device)
            kw = dict(idx_w=[idx]) if axis == 3 else dict(idx_h=[idx])
            x = rt.latent_blend([x], h, w, _lib.BLEND_COPY, **kw)
    frac = float(np.float32(lod) - np.floor(np.float32(lod)))
    x_fade = x
    if frac != 0.0:
        y = _upscale_image(rt, _pool_image(rt, x, 2), 2)
        x_fade = _lerp_lod(rt, x, y, frac)
    factor = int(2 ** int(np.floor(lod)))
    if factor > 1:
        x_orig = _upscale_image(rt, x, factor)
        x_fade = _upscale_image(rt, x_fade, factor) if x_fade is not x else x_orig
    else:
        x_orig = x
    return x_fade, x_orig


class GraphedCritic:
    """One critic's whole loss evaluation `loss.D_wgangp` (forward + backward of the fake and real batches, mixed
    batch, WGAN-GP double backward: ~450 kernel launches) captured ONCE as a CUDA graph and replayed every step.
    Everything in it is shape-static and reads its operands from fixed addresses: the critic's flat variable buffer,
    its flat gradient buffer, and three static input buffers (fake, real, mixing factors) that `__call__` refreshes
    with device copies.  The tensor-core weight planes are re-derived from the variables INSIDE the graph (the
    network's plane cache is emptied before the capture), so a replay always sees the current weights.
    Why: the step is otherwise bound by the host (~45 us of Python per launch, 2 000 launches)."""

    def __init__(self, trainer, name, n):
        self.D, self.grad = trainer.nets[name], trainer.grads[name]
        rt = self.rt = trainer.rt
        dev, res = rt.device, trainer.cfg['resolution']
        self.fake = torch.zeros(n, 3, res, res, dtype=torch.float32, device=dev)
        self.real = torch.zeros(n, 3, res, res, dtype=torch.float32, device=dev)
        self.mix = torch.full((n, 1, 1, 1), 0.5, dtype=torch.float32, device=dev)
        main = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):          # warm-up outside the capture: kernel attributes, allocator pools
            self.fake.uniform_(-1, 1)
            self.real.uniform_(-1, 1)
            self._evaluate(torch.zeros_like(self.grad))
        main.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.D._owner()._prepared.clear()      # capture the weight-plane kernels too
        l0 = rt.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        gc.collect()
        gc_was_on = gc.isenabled()
        gc.disable()                           # finalising unrelated CUDA objects mid-capture would invalidate it
        try:
            with torch.cuda.graph(self.graph):
                self.report = self._evaluate(self.grad)
        finally:
            if gc_was_on:
                gc.enable()
        self.launches = rt.launch_count() - l0
        self.D._owner()._prepared.clear()      # those planes live in the graph's pool; nobody else may keep them

    def _evaluate(self, grad):
        grad.zero_()
        return loss.D_wgangp(self.D, self.fake, self.real, self.mix, grad)

    def __call__(self, fake, real, mix):
        """Refresh the static inputs and replay.  The returned report tensors are the graph's static outputs: read
        them before the next replay."""
        self.fake.copy_(fake)
        self.real.copy_(real)
        self.mix.copy_(mix.reshape(self.mix.shape))
        self.graph.replay()
        return self.report


class GraphedCriticGradient(GraphedCritic):
    """The critic as a fixed function inside the E/G loss (`loss.critic_input_gradient`: forward + data gradients,
    ~150 launches), captured and replayed the same way."""

    def __init__(self, trainer, name, n, weight):
        self.weight = weight
        super().__init__(trainer, name, n)

    def _evaluate(self, grad):
        return loss.critic_input_gradient(self.D, self.fake, self.weight)

    def __call__(self, images):
        self.fake.copy_(images)
        self.graph.replay()
        return self.report


CROP_KEYS = ('eg_crop_interp', 'eg_crop_blend', 'd_interp_crop', 'd_blend_crop')
MIX_KEYS = ('eg_mix', 'd_rec_gp', 'd_interp_gp', 'd_blend_mix', 'd_blend_gp', 'eg_gram_alpha')
IDX_KEYS = ('h_forward', 'w_forward', 'h_backward', 'w_backward')


class _StepGraph:
    """One captured train step: static input buffers, the replay program (CUDA graphs separated by the gradient
    all-reduces when the job has more than one rank) and the static report tensors."""
    __slots__ = ('reals', 'packed', 'mixes', 'draws', 'program', 'report', 'launches', 'pool', 'keep')


class Trainer:
    """Owns the nine networks, the four optimizers and the two gradient buckets (critic phase, E/G phase) of one
    rank.

    Host side of a step (`cuda_graphs='step'`, the default): the FIRST step of a given shape runs eagerly (it also
    loads every kernel), the second one is captured - the whole of run.py:511-513 as one CUDA graph, or three when
    the job has several ranks (cut at the two gradient all-reduces, which stay ordinary NCCL calls) - and every
    later step is: sample the permutations on the host, five small copies into the graph's static buffers, replay.
    What makes the step capturable although `random_crop` moves every step: window SIZES are static, window OFFSETS
    are read by the kernels from device memory (loss.Window, tmx_window_copy)."""

    def __init__(self, config=None, seed=1000, device=None, vgg_weights=None):
        """vgg_weights: the tensorflow_vgg weight dict ({layer: [filter, bias]}, vgg.load_vgg19_npy) or a path to
        vgg19.npy; with it and cfg['loss']['gram_weight'] > 0 (config.py:64: 0.002) the E/G loss includes the VGG-19
        Gram terms.  Without it the term is off (the file is not redistributable: stated deviation)."""
        self.cfg = config or default_config()
        c = self.cfg
        self.rt = Runtime.get(device)
        self.gram = None
        if vgg_weights is not None:
            from .vgg import GramLoss, load_vgg19_npy
            if isinstance(vgg_weights, str):
                vgg_weights = load_vgg19_npy(vgg_weights)
            self.gram = GramLoss(vgg_weights, resolution=c['resolution'], device=self.rt.device)
            c['loss'].setdefault('gram_weight', 0.002)
        res = c['resolution']
        self.nets = {}
        for i, (name, func) in enumerate(NET_FUNCS.items()):                       # run.py:264-269
            self.nets[name] = Network(name, func=func, seed=seed + i, num_channels=3, resolution=res,
                                      device=self.rt.device, **c[name])
        self.G_fcn = Network('G', func=NET_FUNCS['G'], reuse=True, share_vars_with=self.nets['G'], num_channels=3,
                             resolution=res, scale_h=c['scale_h'], scale_w=c['scale_w'], device=self.rt.device,
                             **c['G'])                                              # run.py:273
        parallel.broadcast_([n.flat for n in self.nets.values()])                  # identical replicas
        self.ema = {}
        for src, dst in (('E_zg', 'Es_zg'), ('E_zl', 'Es_zl'), ('G', 'Gs')):       # run.py:270-272
            self.nets[dst] = self.nets[src].clone(dst)
            self.ema[dst] = self.nets[dst].setup_as_moving_average_of(self.nets[src], beta=c['ema_beta'])
        # one flat gradient bucket per optimizer phase = ONE all-reduce per session.run of run.py:511-512
        self.buckets = {'D': GradientBucket({k: self.nets[k] for k in ('D_rec', 'D_interp', 'D_blend')}),
                        'EG': GradientBucket({k: self.nets[k] for k in ('E_zg', 'E_zl', 'G')})}
        self.grads = dict(self.buckets['D'].views)
        self.grads.update(self.buckets['EG'].views)
        self.opts = {}
        for name, members, bucket in (('EG', ('E_zg', 'E_zl', 'G'), 'EG'), ('D_rec', ('D_rec',), 'D'),
                                      ('D_interp', ('D_interp',), 'D'), ('D_blend', ('D_blend',), 'D')):   # run.py:297-300
            opt = Optimizer(name='Train' + name, learning_rate=c['lrate'], **c['opt'])
            for m in members:
                opt.register_gradients(self.nets[m], self.grads[m])
            opt.use_bucket(self.buckets[bucket])
            self.opts[name] = opt
        self.graph_pool = None
        self._critic_graphs = {}
        self._critic_streams = [torch.cuda.Stream(device=self.rt.device) for _ in range(3)]
        self.graph_launches = 0       # kernels replayed from graphs (libtmx counts launches at capture time only)
        self._step_graphs = {}
        self._warm = set()
        self.allreduce_events = []    # (start, end) CUDA events around the phase all-reduces when `time_collectives`
        self.time_collectives = False

    # ------------------------------------------------------------------ host-side random draws of one step
    def sample_draws(self, minibatch, rng, uniform=None):
        """Permutation index vectors (run.py:436-507, same np.random stream as the reference when `uniform` is None)
        plus the graph's own random ops made explicit: one crop offset per loss that crops (loss.py:78-90) and the
        per-sample mixing factors (loss.py:237,329,405,489,505); and, derived from the crop offsets, the crop-aware
        window plans (loss.plan_crop) whose offsets travel to the device with the index vectors."""
        c = self.cfg
        idx = interp.sample_schedule_indices(minibatch, c['latent_res'], c['scale_h'], c['scale_w'], c['levels'], uniform,
                                             block_size=c.get('block_size', 0), perm=c.get('perm', False))
        res, lat = c['resolution'], c['latent_res']
        H, W = lat * c['scale_h'], lat * c['scale_w']
        hi_y, hi_x = res * c['scale_h'] - res, res * c['scale_w'] - res

        def crop():
            return (int(rng.randint(0, hi_y)) if hi_y > 0 else 0, int(rng.randint(0, hi_x)) if hi_x > 0 else 0)

        # every random draw of the step crosses to the device in two page-locked, non-blocking copies: a pageable
        # `.to(device)` per tensor would stall the host behind all queued kernels nine times per step
        mixes = torch.from_numpy(rng.uniform(0.0, 1.0, (len(MIX_KEYS), minibatch, 1, 1, 1)).astype(np.float32))
        crops = {k: crop() for k in CROP_KEYS}
        lod = self.nets['G'].lod
        plans = {k: loss.plan_crop(crops[k], res, lat, H, W, crop_aware=c.get('crop_aware', True), lod=lod)
                 for k in CROP_KEYS}
        offsets = np.array([v for k in CROP_KEYS for v in loss.plan_offsets(plans[k])], np.int32)
        packed = torch.from_numpy(np.concatenate([idx[k].reshape(-1).astype(np.int32) for k in IDX_KEYS] + [offsets]))
        if self.rt.device.type == 'cuda':
            mixes = mixes.pin_memory().to(self.rt.device, non_blocking=True)
            packed = packed.pin_memory().to(self.rt.device, non_blocking=True)
        out = dict(idx=idx, crops=crops, host_plans=plans, _packed=packed, _mixes=mixes)
        out.update(crops)
        out.update(self._draw_views(idx, plans, packed, mixes))
        if self._custom_modes():
            # the tf.random_normal draws of the config-off interpolation modes (loss.py:178,185,187-190, ...): one set
            # per loss graph that samples (D_interp_wgangp, D_blend_wgangp, EG_wgan), forward and reversed branch each
            C = c['E_zl']['latent_channels']
            for phase in ('d_interp', 'd_blend', 'eg'):
                noise = {}
                for which in ('f', 'b'):
                    if c['zg_interp_variational'] == 'variational':
                        noise['zg_' + which] = rng.standard_normal((minibatch, C, 1, 1)).astype(np.float32)
                    if c['zl_interp_variational'] in ('variational', 'random'):
                        noise['zl_' + which] = rng.standard_normal((minibatch, C, H, W)).astype(np.float32)
                out[phase + '_noise'] = {k: torch.from_numpy(v).to(self.rt.device) for k, v in noise.items()}
        return out

    def _custom_modes(self):
        c = self.cfg
        return (c.get('zg_interp_variational', 'hard'), c.get('zl_interp_variational', 'permutational')) != \
            ('hard', 'permutational')

    def _modes(self, draws, phase):
        if not self._custom_modes():
            return None
        c = self.cfg
        return loss.interp_modes(c['zg_interp_variational'], c['zl_interp_variational'], draws.get(phase + '_noise'))

    @staticmethod
    def _draw_views(idx, plans, packed, mixes):
        """Device views of one step's draws inside the `packed` int32 / `mixes` fp32 buffers."""
        idx_dev, off = {}, 0
        for k in IDX_KEYS:
            idx_dev[k] = packed[off:off + idx[k].size].view(idx[k].shape)
            off += idx[k].size
        dev_plans = {}
        for k in CROP_KEYS:
            dev_plans[k] = loss.plan_on_device(plans[k], packed[off:off + 8])
            off += 8
        out = dict(idx_dev=idx_dev, plans=dev_plans)
        out.update({k: mixes[i] for i, k in enumerate(MIX_KEYS)})
        return out

    # ------------------------------------------------------------------ fake images of the canvas critics (no tape)
    def _fcn_fake(self, fwd, which, yx, mix=None, plan=None, modes=None):
        return loss.fcn_fake(self.G_fcn, fwd, which, yx, mix, crop_aware=self.cfg.get('crop_aware', True), plan=plan,
                             modes=modes)

    def _critic(self, name, n):
        key = (name, n, self.nets[name].lod)
        g = self._critic_graphs.get(key)
        if g is None:
            g = self._critic_graphs[key] = GraphedCritic(self, name, n)     # own memory pool: the graphs run concurrently
        self.graph_launches += g.launches
        return g

    def _fork_join(self, jobs):
        """Run the callables of `jobs` on the three side streams, forked from and joined to the current stream.
        The critics are independent of each other: their many small, latency-bound kernels (8x8 / 4x4 maps, dense
        head) overlap instead of queueing behind each other.  Works eagerly and under stream capture."""
        main = torch.cuda.current_stream(self.rt.device)
        fork = torch.cuda.Event()
        fork.record(main)
        results, joins = [], []
        for i, job in enumerate(jobs):
            side = self._critic_streams[i % len(self._critic_streams)]
            side.wait_event(fork)
            with torch.cuda.stream(side):
                results.append(job())
            done = torch.cuda.Event()
            done.record(side)
            joins.append(done)

        def join():
            for done in joins:
                main.wait_event(done)
        return results, join

    def _eg_critic_gradients(self, fwd, draws, use_graphs):
        """The three adversarial terms of the E/G loss (post-step critics as fixed functions) on three streams,
        joined before the generator's backward consumes their image gradients."""
        w = self.cfg['loss']
        jobs = (('rec', 'D_rec', w['rec_G_weight'], lambda: fwd.rec),
                ('interp', 'D_interp', w['interp_G_weight'], lambda: fwd.crop('interp', draws['eg_crop_interp'])),
                ('blend', 'D_blend', w['blend_interp_G_weight'], lambda: fwd.crop('blend', draws['eg_crop_blend'])))
        jobs = [(key, name, wt, img()) for key, name, wt, img in jobs if wt > 0]

        def make(key, name, wt, img):
            if use_graphs:
                gkey = (name, 'input', img.shape[0], self.nets[name].lod, float(wt))
                g = self._critic_graphs.get(gkey)
                if g is None:
                    g = self._critic_graphs[gkey] = GraphedCriticGradient(self, name, img.shape[0], float(wt))
                self.graph_launches += g.launches
                return lambda: g(img)
            return lambda: loss.critic_input_gradient(self.nets[name], img, float(wt))
        results, join = self._fork_join([make(*j) for j in jobs])
        join()
        self._eg_inputs = [j[3] for j in jobs]      # keep the crops alive until the next step (read on side streams)
        return {j[0]: r for j, r in zip(jobs, results)}

    # ------------------------------------------------------------------ one step
    def step(self, reals, draws, lrate=None, phases=('D', 'EG', 'EMA'), lod=None, reals_orig=None, reals_d=None,
             reals_d_orig=None):
        """reals: this rank's share [n,3,R,R] fp32 in [-1,1] on the device (`reals_fade` of run.py:311; `reals_orig`
        - what the encoders see, loss.py:119,126 - defaults to the same tensor, which is exact at integer lod).
        reals_d: the minibatch of the CRITIC phase.  In the reference `reals` is the dataset iterator's get_next()
        inside the graph (run.py:286), so the critic session.run (run.py:511) and the E/G session.run (:512) each
        consume a fresh minibatch: pass both to train like the reference.  None = the critics see the E/G phase's
        minibatch (then one E/G forward serves both phases - 3 % fewer FLOPs, other training dynamics).
        lod: level of detail assigned to all networks before the losses run (run.py:310); None keeps theirs.
        Returns the loss-term report (device scalars; with CUDA graphs they are the graph's static outputs: read
        them before the next step).
        Order of run.py:511-513: critics see the pre-step E/G; E/G see the post-step critics; then EMA."""
        c = self.cfg
        if lod is not None:
            for name in NET_FUNCS:
                self.nets[name].set_lod(lod)
        lod_now = self.nets['G'].lod
        lrate = float(self.opts['EG'].learning_rate if lrate is None else lrate)
        reals_fade = reals
        reals_orig = reals if reals_orig is None else reals_orig
        if reals_d is None:
            d_fade, d_orig = reals_fade, reals_orig
        else:
            d_fade, d_orig = reals_d, (reals_d if reals_d_orig is None else reals_d_orig)
        mode = c.get('cuda_graphs', True)
        mode = 'step' if mode is True else mode
        if os.environ.get('TMX_NO_GRAPH') or lod_now != int(lod_now) or self._custom_modes():
            mode = None      # (a fractional lod changes every step and is baked into the launches; the config-off
            #                   interpolation modes bring per-step noise tensors that have no static graph buffers)
        if os.environ.get('TMX_GRAPH_MODE'):
            mode = os.environ['TMX_GRAPH_MODE']
        if mode == 'step' and 'plans' in draws:
            return self._step_graphed(reals_fade, reals_orig, d_fade, d_orig, draws, lrate, tuple(phases), lod_now)
        return self._step_body(reals_fade, reals_orig, d_fade, d_orig, draws, lrate, phases,
                               critic_graphs=(mode == 'critics'))

    def step_from_dataset(self, images, draws, images_d=None, lod=None, rng=None, drange_data=(0, 255), **step_kwargs):
        """One step from minibatches as the dataset delivers them (run.py:306-312 then 510-513): `images` (and
        `images_d` for the critic phase) are uint8 / float device tensors [n,3,r,r] at the dataset's current resolution
        r = R / 2^floor(lod); `process_reals` applies the dynamic range, the mirror augmentation selected by
        cfg['lr_mirror_augment'] / cfg['ud_mirror_augment'] (config.py:96, run.py:237-238, 311), FadeLOD and
        UpscaleLOD, then `step` runs."""
        c = self.cfg
        lod_now = float(self.nets['G'].lod if lod is None else lod)
        kw = dict(lr_mirror_augment=bool(c.get('lr_mirror_augment', False)),
                  ud_mirror_augment=bool(c.get('ud_mirror_augment', False)), drange_data=drange_data, rng=rng)
        fade, orig = process_reals(images, lod_now, **kw)
        d_fade = d_orig = None
        if images_d is not None:
            d_fade, d_orig = process_reals(images_d, lod_now, **kw)
        return self.step(fade, draws, lod=lod, reals_orig=orig, reals_d=d_fade, reals_d_orig=d_orig, **step_kwargs)

    def _allreduce(self, name, overlap=True):
        """Issue the SUM all-reduce of gradient bucket `name` on the communicator's stream and return wait(): kernels
        launched before wait() overlap the collective (the VGG-19 Gram terms ride on the critics' all-reduce), wait()
        makes the compute stream wait for it.  The recorded events bracket the EXPOSED part only (wait entry -> done)."""
        done = self.buckets[name].allreduce_async()

        def wait():
            if self.time_collectives and parallel.world_size() > 1:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                done()
                e1.record()
                self.allreduce_events.append((name, e0, e1))
            else:
                done()
        return wait

    def _step_body(self, reals_fade, reals_orig, d_fade, d_orig, draws, lrate, phases, critic_graphs=False,
                   boundary=None):
        """The launches of one step.  `boundary(name)`: where the all-reduce of gradient bucket `name` is ISSUED; it
        returns wait(), called where the reduced gradients are needed (default: Trainer._allreduce; the graph capture
        cuts the step at both points)."""
        boundary = boundary or self._allreduce
        report = {}
        c = self.cfg
        nets = self.nets
        n = reals_orig.shape[0]
        idx = draws.get('idx_dev', draws['idx'])
        plans = draws.get('plans')
        if plans is None:        # draws made by hand: host plans
            res, lat = c['resolution'], c['latent_res']
            plans = {k: loss.plan_crop(draws[k], res, lat, lat * c['scale_h'], lat * c['scale_w'],
                                       crop_aware=c.get('crop_aware', True), lod=nets['G'].lod) for k in CROP_KEYS}
        shared = d_orig is reals_orig and d_fade is reals_fade

        def eg_forward():
            return loss.EGForward(nets['E_zg'], nets['E_zl'], nets['G'], self.G_fcn, reals_orig, idx, draws['eg_mix'],
                                  c['scale_h'], c['scale_w'], defer_canvases=True,
                                  plans={'interp': plans['eg_crop_interp'], 'blend': plans['eg_crop_blend']},
                                  modes=self._modes(draws, 'eg'))
        fwd = gram_grads = None
        if 'D' in phases:
            if shared:
                fwd = fwd_d = eg_forward()
            else:        # the critics' own minibatch: encoders + reconstruction without tapes (loss.py:308-320)
                fwd_d = loss.EGForward(nets['E_zg'], nets['E_zl'], nets['G'], self.G_fcn, d_orig, idx, draws['eg_mix'],
                                       c['scale_h'], c['scale_w'], defer_canvases=True, record=False,
                                       plans={'interp': plans['d_interp_crop'], 'blend': plans['d_blend_crop']})
            fakes = (('D_rec', fwd_d.rec, 'd_rec_gp'),
                     ('D_interp', self._fcn_fake(fwd_d, 'interp', draws['d_interp_crop'], plan=plans['d_interp_crop'],
                                                 modes=self._modes(draws, 'd_interp')), 'd_interp_gp'),
                     ('D_blend', self._fcn_fake(fwd_d, 'blend', draws['d_blend_crop'], draws['d_blend_mix'],
                                                plan=plans['d_blend_crop'], modes=self._modes(draws, 'd_blend')),
                      'd_blend_gp'))
            if critic_graphs:
                self.buckets['D'].marks.zero_()      # (each critic graph zeroes its own gradient view)
            else:
                self.buckets['D'].zero_()

            def critic_job(name, fake, gp):
                if critic_graphs:
                    g = self._critic(name, n)            # captured on the current stream the first time
                    return lambda: g(fake, d_fade, draws[gp])
                return lambda: loss.D_wgangp(nets[name], fake, d_fade, draws[gp], self.grads[name])
            jobs = [critic_job(*f) for f in fakes]
            if os.environ.get('TMX_NO_FORK'):
                results, join = [j() for j in jobs], (lambda: None)
            else:
                results, join = self._fork_join(jobs)
            # the taped evaluations of the E/G phase do not depend on the critics: run them on the main stream WHILE
            # the critics (many thin, low-occupancy kernels) run on the side streams
            if 'EG' in phases:
                if fwd is None:
                    fwd = eg_forward()
                fwd.decode_canvases()
            join()                   # fakes / reals are only released after this join
            for (name, _, _), rep in zip(fakes, results):
                report.update({name + '/' + k: v for k, v in rep.items()})
            for name in ('D_rec', 'D_interp', 'D_blend'):                           # one session.run (run.py:511)
                self.opts[name].mark()
            wait_d = boundary('D')
            # in flight now: the critics' gradient all-reduce.  The VGG-19 Gram terms of the E/G loss need the generated
            # images and the reals, not the critics: they run while the collective does
            if 'EG' in phases and self.gram is not None and c['loss'].get('gram_weight', 0.0) > 0:
                lw = c['loss']
                gram_grads = loss.gram_terms(fwd, draws['eg_crop_interp'], draws['eg_crop_blend'], self.gram,
                                             lw['gram_weight'], draws.get('eg_gram_alpha'),
                                             lw.get('interp_G_weight', 1.0), lw.get('blend_interp_G_weight', 1.0),
                                             reals_fade=reals_fade)
            wait_d()
            for name in ('D_rec', 'D_interp', 'D_blend'):
                report[name + '/skipped'] = self.opts[name].update(lrate)
            self._d_keep = fakes
        if 'EG' in phases:
            if fwd is None:
                fwd = eg_forward()
            fwd.decode_canvases()
            self.buckets['EG'].zero_()
            critic_grads = None
            if not os.environ.get('TMX_NO_FORK'):
                critic_grads = self._eg_critic_gradients(fwd, draws, critic_graphs)
            rep = loss.EG_backward(fwd, nets['D_rec'], nets['D_interp'], nets['D_blend'],
                                   draws['eg_crop_interp'], draws['eg_crop_blend'], self.grads, reals_fade=reals_fade,
                                   critic_grads=critic_grads, gram=self.gram, gram_alpha=draws.get('eg_gram_alpha'),
                                   gram_grads=gram_grads, **c['loss'])
            report.update({'EG/' + k: v for k, v in rep.items()})
            self.opts['EG'].mark()
            boundary('EG', overlap=False)()         # (nothing of this step is left to overlap with the E/G all-reduce)
            report['EG/skipped'] = self.opts['EG'].update(lrate)
        del fwd
        if 'EMA' in phases:
            for upd in self.ema.values():
                upd()
        return report

    # ------------------------------------------------------------------ the step as CUDA graphs
    def _step_graphed(self, reals_fade, reals_orig, d_fade, d_orig, draws, lrate, phases, lod_now):
        shared = d_orig is reals_orig and d_fade is reals_fade
        distinct = []        # the distinct input tensors, in a fixed role order
        roles = []
        for t in (reals_fade, reals_orig, d_fade, d_orig):
            for i, u in enumerate(distinct):
                if u is t:
                    roles.append(i)
                    break
            else:
                roles.append(len(distinct))
                distinct.append(t)
        sig = tuple(None if draws['host_plans'][k][w] is None else tuple(draws['host_plans'][k][w][2:])
                    for k in CROP_KEYS for w in loss.PLAN_KEYS)
        key = (tuple(reals_fade.shape), tuple(roles), lod_now, lrate, phases, sig, parallel.world_size())
        ent = self._step_graphs.get(key)
        if ent is None:
            if key not in self._warm:
                # first step of this shape: eager (loads the kernels, sizes the allocator pools, creates the
                # optimizers' device state) - and it is a real step
                self._warm.add(key)
                return self._step_body(reals_fade, reals_orig, d_fade, d_orig, draws, lrate, phases)
            ent = self._step_graphs[key] = self._capture_step(distinct, roles, draws, lrate, phases)
        for dst, src in zip(ent.reals, distinct):
            dst.copy_(src, non_blocking=True)
        ent.packed.copy_(draws['_packed'], non_blocking=True)
        ent.mixes.copy_(draws['_mixes'], non_blocking=True)
        pending = {}
        for kind, obj in ent.program:
            if kind == 'graph':
                obj.replay()
            elif kind == 'allreduce':
                pending[obj] = self._allreduce(obj)
            else:                                # 'wait'
                pending.pop(obj)()
        self.graph_launches += ent.launches
        for net in self.nets.values():      # the graph re-derives its weight planes itself; eager users must too
            net._touch()
        return ent.report

    def _capture_step(self, distinct, roles, draws, lrate, phases):
        rt = self.rt
        dev = rt.device
        ent = _StepGraph()
        ent.reals = [torch.empty_like(t) for t in distinct]
        ent.packed = torch.empty_like(draws['_packed'])
        ent.mixes = torch.empty_like(draws['_mixes'])
        for dst, src in zip(ent.reals, distinct):
            dst.copy_(src)
        ent.packed.copy_(draws['_packed'])
        ent.mixes.copy_(draws['_mixes'])
        sdraws = dict(draws)
        sdraws.update(self._draw_views(draws['idx'], draws['host_plans'], ent.packed, ent.mixes))
        ent.draws = sdraws
        x = [ent.reals[i] for i in roles]
        torch.cuda.synchronize(dev)
        for net in list(self.nets.values()) + [self.G_fcn]:
            net._owner()._prepared.clear()       # capture the weight-plane kernels too
        ent.pool = torch.cuda.graph_pool_handle()
        cap_stream = torch.cuda.Stream(device=dev)
        program, state = [], {}

        def begin():
            g = torch.cuda.CUDAGraph()
            ctx = torch.cuda.graph(g, pool=ent.pool, stream=cap_stream)
            ctx.__enter__()
            state['g'], state['ctx'] = g, ctx

        def end():
            state['ctx'].__exit__(None, None, None)
            program.append(('graph', state['g']))

        def boundary(name, overlap=True):
            if parallel.world_size() == 1:
                return lambda: None
            end()                                # the all-reduce stays an ordinary NCCL call between two graphs
            program.append(('allreduce', name))
            if not overlap:                      # waited for at once: no (empty) graph in between
                program.append(('wait', name))
                begin()
                return lambda: None
            begin()

            def wait():                          # ... and so does the point where the compute stream waits for it
                end()
                program.append(('wait', name))
                begin()
            return wait
        l0 = rt.launch_count()
        gc.collect()
        gc_was_on = gc.isenabled()
        gc.disable()                           # finalising unrelated CUDA objects mid-capture would invalidate it
        try:
            begin()
            try:
                ent.report = self._step_body(x[0], x[1], x[2], x[3], sdraws, lrate, phases, boundary=boundary)
                ent.keep = (getattr(self, '_d_keep', None), getattr(self, '_eg_inputs', None))
            finally:
                end()
        finally:
            if gc_was_on:
                gc.enable()
        ent.launches = rt.launch_count() - l0
        ent.program = program
        for net in list(self.nets.values()) + [self.G_fcn]:
            net._owner()._prepared.clear()       # those planes live in the graph's pool; nobody else may keep them
        return ent
