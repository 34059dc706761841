"""Central value (asymmetric critic) on top of the continuous B200 agent -- SURVEY.md 8f rank 1 (reference:
algos_torch/central_value.py CentralValueTrain, models.py:425-464 ModelCentralValue, a2c_common.py:250-262, :593-615, :1010-1011,
:1536-1537, :1651-1660; a2c_continuous.py:50-75).

STATUS: parity-green on a B200 (round 2, ``tests/test_cv_gpu.py``).  The reference
semantics are pinned by ``oracle.ppo_oracle.CentralValueOracle`` against ``tests/golden/agent_cv.pt``; the HOST logic of this module runs
on CPU against the same golden vectors with torch stand-ins for the kernels (``tests/test_agent_cv_host_cpu.py``); the only new kernel
(``csrc/critic.cu`` value loss) has its row arithmetic exercised on the CPU too.  Everything else reuses validated kernels.

``A2CAgentCV`` subclasses the validated ``A2CAgent`` without touching it:
  * the env returns ``{'obs': actor view, 'states': privileged critic view}``; ``states`` go to the critic's arena;
  * rollout values / the last value come from the critic (de-normalised with ITS value normaliser), so the time-out bootstrap and
    GAE use them; the agent's value normaliser IS the critic's (a2c_continuous.py:72-73);
  * the critic trains (own minibatching, own Adam, own obs normaliser re-armed on every minibatch) before the actor's mini-epochs;
  * the actor keeps its own value head and value loss (``use_experimental_cv`` defaults to True in the reference).
Scope of this first edition: one GPU, MLP critic, fp32 kernels for the critic, no whole-epoch CUDA graph (the update-phase graph of the
actor still works), identity / linear critic LR schedule.
"""
from collections import OrderedDict

import torch
import torch.distributed as dist

from . import ops
from .agent import A2CAgent
from .common import IdentityScheduler, LinearScheduler
from .model import _RunningStats, check_network_params


class CentralValueNet:
    """CentralValueTrain on the fp32 kernels: flat arena [W0 b0 ... W_v b_v] in the reference's parameter order."""

    # `CentralValueTrain.model` (central_value.py:44): the reference's Runner wraps it with torch.compile unless the YAML says
    # `torch_compile: False` (torch_runner.py:308-313).  Here the net IS its model and there is no nn.Module to compile: reading `.model`
    # gives the net, assigning anything else to it is ignored with a note (like CompileTolerantModel does for `agent.model`).
    @property
    def model(self):
        return self

    @model.setter
    def model(self, m):
        if m is not self and not self.__dict__.get('_b200_compile_noted'):
            print('b200: torch.compile of central_value_net.model ignored (no nn.Module on this path; set torch_compile: False to silence)')
            self.__dict__['_b200_compile_noted'] = True

    def __call__(self, *args, **kwargs):
        raise NotImplementedError('the central value net is a flat parameter arena driven by kernels (get_value / train_net); '
                                  'it has no module-style forward')

    def __init__(self, cv_config, state_dim, num_actors, horizon, normalize_value, max_epochs, device, multi_gpu=False, world_size=1):
        net = cv_config['network']
        self.multi_gpu, self.world_size = bool(multi_gpu) and world_size > 1, int(world_size)
        self.writter = None             # the agent's summary writer (central_value.py:79; the reference's spelling)
        check_network_params(net)
        mlp = net['mlp']
        self.units = list(mlp['units'])
        self.activation = mlp.get('activation', 'elu')
        if self.activation not in ops.ACT or not self.units or any(k in net for k in ('cnn', 'rnn')):
            raise NotImplementedError('central value network: a plain MLP is supported')
        self.act_id = ops.ACT[self.activation]
        self.S, self.N, self.H = int(state_dim), int(num_actors), int(horizon)
        self.device = torch.device(device)
        self.normalize_input, self.normalize_value = bool(cv_config['normalize_input']), bool(normalize_value)
        self.lr = float(cv_config['learning_rate'])
        self.mini_epoch = int(cv_config['mini_epochs'])
        self.batch_size = self.N * self.H
        if 'minibatch_size' not in cv_config and 'minibatch_size_per_env' not in cv_config:
            raise ValueError("Configuration must include either 'minibatch_size' or 'minibatch_size_per_env'. "
                             "Neither was found in the provided config.")
        self.minibatch_size = int(cv_config.get('minibatch_size', self.N * cv_config.get('minibatch_size_per_env', 0)))
        if self.minibatch_size <= 0 or self.batch_size % self.minibatch_size or self.minibatch_size % self.H:
            raise NotImplementedError('central value minibatch_size must divide the batch and be a multiple of horizon_length')
        self.num_minibatches = self.batch_size // self.minibatch_size
        self.envs_per_mb = self.minibatch_size // self.H
        self.clip_value, self.e_clip = bool(cv_config['clip_value']), float(cv_config.get('e_clip', 0.2))
        self.truncate_grads, self.grad_norm = bool(cv_config.get('truncate_grads', False)), float(cv_config.get('grad_norm', 1))
        self.weight_decay = float(cv_config.get('weight_decay', 0.0))
        self.freeze = bool(cv_config.get('freeze_critic', False))
        self.scheduler = (LinearScheduler(self.lr, max_steps=max_epochs, apply_to_entropy=False, start_entropy_coef=0)
                          if cv_config.get('lr_schedule') == 'linear' else IdentityScheduler())        # central_value.py:55-62
        self.epoch_num = self.frame = 0
        sizes, ins = [], self.S
        for i, u in enumerate(self.units):
            sizes += [(f'W{i}', (u, ins)), (f'b{i}', (u,))]
            ins = u
        self.Hl = ins
        sizes += [('W_v', (1, ins)), ('b_v', (1,))]
        self.layout, off = OrderedDict(), 0
        for n, shp in sizes:
            self.layout[n] = (off, shp)
            off += int(torch.Size(shp).numel())
        self.num_params = off
        dev = self.device
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        for n, (o, shp) in self.layout.items():
            if n.startswith('W'):
                bound = 1.0 / float(shp[1]) ** 0.5
                self.view(n).uniform_(-bound, bound)
        self.running_mean_std = _RunningStats(self.S, dev) if self.normalize_input else None
        self.value_mean_std = _RunningStats(1, dev) if self.normalize_value else None
        f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)   # noqa: E731
        mb = self.minibatch_size
        self.states = f(self.H, self.N, self.S)
        self.r_act = [f(self.N, u) for u in self.units]
        self.t_act = [f(mb, u) for u in self.units]
        self.d_act = [f(mb, u) for u in self.units]
        self.r_val, self.t_val, self.d_val = f(self.N, 1), f(mb, 1), f(mb, 1)
        self.n_splits = max(1, min(64, mb // 256))
        self.part = f(self.n_splits, off)
        self.grad = f(off)
        self.loss_partials = torch.zeros((mb + 255) // 256, 8, dtype=torch.float64, device=dev)
        self.inv_counts = f(self.num_minibatches)
        self.mom_scratch = torch.zeros(148 * 4 * 2 * self.S, dtype=torch.float64, device=dev)
        self.counters = torch.zeros(4, dtype=torch.int32, device=dev)
        self.opt_state = torch.tensor([self.lr, 0.0, 0.0, 0.0], dtype=torch.float64, device=dev)
        self.adam_stats = f(16)
        # grad_scale = 1/world: the optimiser kernel scales the SUM the all-reduce leaves in self.grad (central_value.py:322-337)
        self.opt_cfg = ops.OptCfg(0.9, 0.999, 1e-8, self.weight_decay, self.grad_norm, 0.0, 1e-6, 1e-2, 1.5,
                                  1.0 / (self.world_size if self.multi_gpu else 1), int(self.truncate_grads), 0)
        self.last_losses = []

    def view(self, name, arena=None):
        o, shp = self.layout[name]
        a = self.flat if arena is None else arena
        return a[o:o + int(torch.Size(shp).numel())].view(shp)

    def param_names(self):
        names = []
        for i in range(len(self.units)):
            names += [f'a2c_network.actor_mlp.{2 * i}.weight', f'a2c_network.actor_mlp.{2 * i}.bias']
        return names + ['a2c_network.value.weight', 'a2c_network.value.bias']

    def _param_views(self, arena=None):
        out = []
        for i in range(len(self.units)):
            out += [self.view(f'W{i}', arena), self.view(f'b{i}', arena)]
        return out + [self.view('W_v', arena), self.view('b_v', arena)]

    def state_dict(self):
        sd = OrderedDict()
        if self.normalize_value:
            sd.update(self.value_mean_std.state_dict('value_mean_std.'))
        if self.normalize_input:
            sd.update(self.running_mean_std.state_dict('running_mean_std.'))
        for n, v in zip(self.param_names(), self._param_views()):
            sd[n] = v.clone()
        return sd

    def load_state_dict(self, sd, strict=True):
        def strip(k):          # CentralValueTrain.state_dict() prefixes its model's keys with 'model.'; torch.compile adds '_orig_mod.'
            k = k.replace('_orig_mod.', '')
            return k[len('model.'):] if k.startswith('model.') else k
        sd = {strip(k): v for k, v in sd.items()}
        with torch.no_grad():
            for n, v in zip(self.param_names(), self._param_views()):
                if n in sd:
                    v.copy_(sd[n].reshape(v.shape))
                elif strict:
                    raise KeyError(f'missing key in central value state_dict: {n}')
            if self.normalize_value and 'value_mean_std.running_mean' in sd:
                self.value_mean_std.load_state_dict(sd, 'value_mean_std.')
            if self.normalize_input and 'running_mean_std.running_mean' in sd:
                self.running_mean_std.load_state_dict(sd, 'running_mean_std.')

    # ---------------------------------------------------------------------------------------------- forward
    def _trunk(self, x, acts, M, rows_per_chunk=None, chunk_stride=0):
        r = self.running_mean_std
        nm, ns = (r.mean_f32, r.std_f32) if self.normalize_input else (None, None)
        ops.linear_fwd(x, self.view('W0'), self.view('b0'), acts[0], self.act_id, rows_per_chunk=rows_per_chunk, chunk_stride=chunk_stride,
                       x_ld=self.S, norm_mean=nm, norm_std=ns, M=M)
        for i in range(1, len(self.units)):
            ops.linear_fwd(acts[i - 1], self.view(f'W{i}'), self.view(f'b{i}'), acts[i], self.act_id, M=M)

    def get_value(self, states, out):
        """central_value.py:207-229 (eval mode): out[N] = denorm(critic(states))"""
        self._trunk(states, self.r_act, self.N)
        ops.linear_fwd(self.r_act[-1], self.view('W_v'), self.view('b_v'), self.r_val, 0, M=self.N)
        if self.normalize_value:
            vm = self.value_mean_std
            ops.normalize(self.r_val, vm.running_mean, vm.running_var, denorm=True, out=out.view(self.N, 1))
        else:
            out.copy_(self.r_val.view(-1))

    # ---------------------------------------------------------------------------------------------- training (central_value.py:246-274)
    def train_net(self, old_values_n, returns_n, valid):
        H, N, mb, epm, P, S_ = self.H, self.N, self.minibatch_size, self.envs_per_mb, self.num_params, self.n_splits
        if valid is not None:
            ops.mask_inv_counts(valid, H, N, epm, self.inv_counts)
        rows = []
        for _ in range(self.mini_epoch):
            if self.freeze:
                break
            for i in range(self.num_minibatches):
                e0 = i * epm
                x = self.states[0, e0:]
                if self.normalize_input:      # train_critic -> self.train() re-arms the critic's obs normaliser on every minibatch
                    r = self.running_mean_std
                    ops.moments_update(x, self.S, epm, H, N, r.running_mean, r.running_var, r.count, r.mean_f32, r.std_f32, self.mom_scratch,
                                       self.counters[1:2])
                self._trunk(x, self.t_act, mb, rows_per_chunk=epm, chunk_stride=N)
                ops.linear_fwd(self.t_act[-1], self.view('W_v'), self.view('b_v'), self.t_val, 0, M=mb)
                nb = ops.value_loss(self.t_val, 1, old_values_n[0, e0:], returns_n[0, e0:], None if valid is None else valid[0, e0:], epm, N, mb,
                                    self.e_clip, self.clip_value, None if valid is None else self.inv_counts[i:i + 1], self.d_val, 1,
                                    self.loss_partials)
                rows.append(self.loss_partials[:nb, 0].sum().float())
                off_w, _ = self.layout['W_v']
                off_b, _ = self.layout['b_v']
                ops.linear_bwd_weight(self.d_val, self.t_act[-1], self.part[0, off_w:], self.part[0, off_b:], self.Hl, 1, S_, M=mb, split_stride=P)
                ops.linear_bwd_data(self.d_val, self.view('W_v'), self.t_act[-1], self.d_act[-1], self.act_id, M=mb)
                r = self.running_mean_std
                nm, ns = (r.mean_f32, r.std_f32) if self.normalize_input else (None, None)
                for l in range(len(self.units) - 1, -1, -1):
                    off_w, shp = self.layout[f'W{l}']
                    off_b, _ = self.layout[f'b{l}']
                    if l > 0:
                        ops.linear_bwd_weight(self.d_act[l], self.t_act[l - 1], self.part[0, off_w:], self.part[0, off_b:], shp[1], shp[0], S_,
                                              M=mb, split_stride=P)
                        ops.linear_bwd_data(self.d_act[l], self.view(f'W{l}'), self.t_act[l - 1], self.d_act[l - 1], self.act_id, M=mb)
                    else:
                        ops.linear_bwd_weight(self.d_act[0], x, self.part[0, off_w:], self.part[0, off_b:], shp[1], shp[0], S_,
                                              rows_per_chunk=epm, chunk_stride=N, x_ld=self.S, norm_mean=nm, norm_std=ns, M=mb, split_stride=P)
                ops.reduce_splits(self.part, self.grad, P, S_, split_stride=P)
                if self.multi_gpu:      # the critic's own flat-gradient exchange, one per minibatch (central_value.py:322-337)
                    dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)
                ops.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.opt_state, None, self.opt_cfg, self.adam_stats,
                              self.counters[2:3], n=P)
        self.last_losses = rows
        self.epoch_num += 1
        self.lr, _ = self.scheduler.update(self.lr, 0, self.epoch_num, self.frame, 0)
        self.opt_state[0] = self.lr
        self.frame += self.batch_size
        if self.writter is not None and rows:       # central_value.py:262-272 (one read-back per epoch, like the reference's add_scalar)
            self.writter.add_scalar('losses/cval_loss', float(torch.stack(rows).sum()) / (self.mini_epoch * self.num_minibatches), self.frame)
            self.writter.add_scalar('info/cval_lr', self.lr, self.frame)
        return rows


class A2CAgentCV(A2CAgent):
    def __init__(self, base_name, params):
        config = params['config']
        cv_config = config.get('central_value_config')
        if cv_config is None:
            raise ValueError('A2CAgentCV needs central_value_config')
        config['central_value_config'] = None          # the base class refuses it; everything it builds is unchanged by the critic
        try:
            super().__init__(base_name, params)
        finally:
            config['central_value_config'] = cv_config
        self.central_value_config = cv_config
        self.has_central_value = True
        self.use_experimental_cv = config.get('use_experimental_cv', True)
        if not self.use_experimental_cv:
            raise NotImplementedError('use_experimental_cv: False (actor without a value loss)')
        space = self.env_info.get('state_space', None) or self.observation_space           # a2c_common.py:254-262
        if type(space).__name__ == 'Dict' or len(space.shape) != 1:
            raise NotImplementedError('only flat state spaces')
        self.state_shape = space.shape
        self.central_value_net = CentralValueNet(cv_config, space.shape[0], self.num_actors, self.horizon_length, self.normalize_value,
                                                 self.max_epochs, self.device_t, multi_gpu=self.multi_gpu, world_size=self.world_size)
        self.central_value_net.writter = self.writer
        self.value_mean_std = self.central_value_net.value_mean_std                        # a2c_continuous.py:72-73
        self._states = None

    # ---- env plumbing: keep the privileged view
    def obs_to_tensors(self, obs):
        if isinstance(obs, dict) and 'states' in obs:
            self._states = self.cast_obs(obs['states'], name='states')
        elif not isinstance(obs, dict) and self.env_info.get('state_space', None) is None:
            self._states = self.cast_obs(obs, name='states')          # state_space fallback = observation_space
        return super().obs_to_tensors(obs)

    # ---- rollout: the critic supplies the values (a2c_common.py:593-600, :603-615, :1010-1011)
    def _policy_step(self, obs, t, noise=None):
        super()._policy_step(obs, t, noise)
        cv = self.central_value_net
        cv.states[t].copy_(self._states)
        cv.get_value(self._states, self.values[t])

    def get_values(self, obs):
        self.central_value_net.get_value(self._states, self.last_values)

    def _whole_epoch_graph_ok(self):
        return False                                   # the critic's training is eager

    def _stats_modules(self):
        """a2c_common.py:753-765: the cross-rank running-stat sync also covers the critic's normalisers (the value normaliser the
        agent uses IS the critic's; the actor model's own never moves and merges as a no-op delta)"""
        mods = super()._stats_modules()
        cv = self.central_value_net
        if cv.running_mean_std is not None:
            mods.append(('cv_obs', cv.running_mean_std))
        if cv.value_mean_std is not None:
            mods.append(('cv_value', cv.value_mean_std))
        return mods

    # ---- the agent's value normaliser is the critic's
    def _prepare(self, n_partials):
        m, keep = self.model, self.model.value_mean_std
        m.value_mean_std = self.central_value_net.value_mean_std
        try:
            super()._prepare(n_partials)
        finally:
            m.value_mean_std = keep

    # ---- the critic trains before the actor's mini-epochs (a2c_common.py:1536-1537)
    def _run_update(self):
        self.central_value_net.train_net(self.old_values_n, self.returns_n, self.valid)
        super()._run_update()

    def train_central_value(self):
        return self.central_value_net.train_net(self.old_values_n, self.returns_n, self.valid)

    # ---- checkpoints (a2c_common.py:831-833, :859-862)
    def get_full_state_weights(self):
        state = super().get_full_state_weights()
        cv = self.central_value_net
        state['assymetric_vf_nets'] = OrderedDict(('model.' + k, v) for k, v in cv.state_dict().items())
        st = {i: {'step': torch.tensor(float(cv.opt_state.cpu()[1])), 'exp_avg': m.clone(), 'exp_avg_sq': v.clone()}
              for i, (m, v) in enumerate(zip(cv._param_views(cv.exp_avg), cv._param_views(cv.exp_avg_sq)))}
        state['assymetric_vf_optimizer'] = {'state': st, 'param_groups': [{'lr': cv.lr, 'betas': (0.9, 0.999), 'eps': 1e-08,
                                                                          'weight_decay': cv.weight_decay, 'params': list(range(len(st)))}]}
        return state

    def restore_central_value_function(self, fn):
        """a2c_continuous.py:90-92: the critic's weights and normalisers only (`load_critic_only` of the Runner, torch_runner.py:43-50)"""
        checkpoint = torch.load(fn, map_location=self.device_t, weights_only=False)
        self.set_central_value_function_weights(checkpoint)

    def set_central_value_function_weights(self, weights):
        """a2c_common.py:885-887"""
        self.central_value_net.load_state_dict(weights['assymetric_vf_nets'])
        self._seed_stats_sync_snapshots()

    def set_full_state_weights(self, weights, set_epoch=True):
        super().set_full_state_weights(weights, set_epoch=set_epoch)
        cv = self.central_value_net
        if 'assymetric_vf_nets' in weights:
            cv.load_state_dict(weights['assymetric_vf_nets'])
        osd = weights.get('assymetric_vf_optimizer')
        if osd:
            step = 0
            for i, (m, v) in enumerate(zip(cv._param_views(cv.exp_avg), cv._param_views(cv.exp_avg_sq))):
                if i in osd['state']:
                    m.copy_(osd['state'][i]['exp_avg'].reshape(m.shape))
                    v.copy_(osd['state'][i]['exp_avg_sq'].reshape(v.shape))
                    step = int(float(osd['state'][i]['step']))
            cv.lr = float(osd['param_groups'][0]['lr'])
            cv.opt_state.copy_(torch.tensor([cv.lr, float(step), 0.9 ** step if step else 0.0, 0.999 ** step if step else 0.0],
                                            dtype=torch.float64))
