"""CPU oracle for the PPO hot path (rollout -> GAE -> minibatch update).

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import this module.
The product path (``rl_games_b200``) never routes through it and fails loudly when
its CUDA library is missing.

This is a plain-PyTorch (CPU, fp32 tensors + fp64 running statistics) restatement of
the reference algorithm, function by function; every function cites the reference
``file:line`` (relative to Denys88/rl_games @ 262cf20) that it follows.  Parity is
PINNED: ``tests/golden/gen_golden.py`` imports the real reference in the build
container and stores its outputs on seeded inputs under ``tests/golden/*.pt``;
``tests/test_oracle_vs_golden.py`` checks this file against those vectors.
"""
import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# GAE -- rl_games/triton_kernels/gae_kernel.py:62-79 (_pytorch_gae), semantics of :16-59
# ----------------------------------------------------------------------------------------------
def gae(mb_rewards, mb_values, mb_dones, last_values, last_dones, gamma, tau):
    """mb_rewards/mb_values [H,N,V]; mb_dones [H,N] float; last_values [N,V]; last_dones [N]."""
    horizon = mb_rewards.shape[0]
    mb_advs = torch.zeros_like(mb_rewards)
    lastgaelam = 0
    for t in reversed(range(horizon)):
        if t == horizon - 1:
            nextnonterminal = 1.0 - last_dones
            nextvalues = last_values
        else:
            nextnonterminal = 1.0 - mb_dones[t + 1]
            nextvalues = mb_values[t + 1]
        nextnonterminal = nextnonterminal.unsqueeze(1)
        delta = mb_rewards[t] + gamma * nextvalues * nextnonterminal - mb_values[t]
        mb_advs[t] = lastgaelam = delta + gamma * tau * nextnonterminal * lastgaelam
    return mb_advs


def gae_f64_scalar(mb_rewards, mb_values, mb_dones, last_values, last_dones, gamma, tau):
    """Independent fp64 scalar recursion -- tests/test_triton_gae.py:20-42 (reference_gae)."""
    horizon, num_envs, value_size = mb_rewards.shape
    r, v, d = mb_rewards.double(), mb_values.double(), mb_dones.double()
    lv, ld = last_values.double(), last_dones.double()
    advs = torch.zeros_like(r)
    last = torch.zeros(num_envs, value_size, dtype=torch.float64)
    for t in reversed(range(horizon)):
        if t == horizon - 1:
            nv, nnt = lv, (1.0 - ld)
        else:
            nv, nnt = v[t + 1], (1.0 - d[t + 1])
        nnt = nnt.unsqueeze(1)
        delta = r[t] + gamma * nv * nnt - v[t]
        last = delta + gamma * tau * nnt * last
        advs[t] = last
    return advs


# ----------------------------------------------------------------------------------------------
# swap_and_flatten01 -- rl_games/common/a2c_common.py:33-40.  flat index = env*H + t
# ----------------------------------------------------------------------------------------------
def swap_and_flatten01(arr):
    if arr is None:
        return arr
    s = arr.size()
    return arr.transpose(0, 1).reshape(s[0] * s[1], *s[2:])


# ----------------------------------------------------------------------------------------------
# masked moments -- rl_games/algos_torch/torch_ext.py:157-191
# ----------------------------------------------------------------------------------------------
def get_mean_var_with_masks(values, masks):
    sum_mask = masks.sum().clamp(min=1.0)
    values_mask = values * masks
    values_mean = values_mask.sum() / sum_mask
    min_sqr = ((((values_mask) ** 2) / sum_mask).sum() - ((values_mask / sum_mask).sum()) ** 2)
    values_var = min_sqr * sum_mask / (sum_mask - 1).clamp(min=1.0)
    return values_mean, values_var


def normalization_with_masks(values, masks):
    if masks is None:
        return (values - values.mean()) / (values.std() + 1e-8)
    values_mean, values_var = get_mean_var_with_masks(values, masks)
    values_std = torch.sqrt(values_var)
    return (values - values_mean) / (values_std + 1e-8)


def apply_masks(losses, mask=None):
    sum_mask = None
    if mask is not None:
        mask = mask.unsqueeze(1)
        sum_mask = mask.sum().clamp(min=1.0)
        res_losses = [(l * mask).sum() / sum_mask for l in losses]
    else:
        res_losses = [torch.mean(l) for l in losses]
    return res_losses, sum_mask


# ----------------------------------------------------------------------------------------------
# RunningMeanStd -- rl_games/algos_torch/running_mean_std.py:19-114
# fp64 mean/var, int64 count; init mean 0, var 1, count 1; Chan merge :55-67
# ----------------------------------------------------------------------------------------------
class RunningMeanStd:
    def __init__(self, insize, epsilon=1e-5):
        self.insize = insize
        self.epsilon = epsilon
        self.running_mean = torch.zeros(insize, dtype=torch.float64)
        self.running_var = torch.ones(insize, dtype=torch.float64)
        self.count = torch.ones((), dtype=torch.int64)
        self.training = False

    def train(self):
        self.training = True

    def eval(self):
        self.training = False

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        count_f = self.count.to(self.running_mean.dtype)
        tot = count_f + batch_count
        delta = batch_mean - self.running_mean
        new_mean = self.running_mean + delta * batch_count / tot
        m_a = self.running_var * count_f
        m_b = batch_var * batch_count
        M2 = m_a + m_b + delta ** 2 * count_f * batch_count / tot
        self.running_mean = new_mean
        self.running_var = M2 / tot
        self.count = self.count + batch_count

    def __call__(self, x, denorm=False, mask=None):
        if self.training:
            if mask is not None:
                mean, var = get_mean_var_with_masks(x, mask)
            else:
                mean = x.mean(0)
                var = x.var(0, unbiased=False)
            self.update_from_moments(mean, var, x.size(0))
        cm, cv = self.running_mean, self.running_var
        if denorm:
            y = torch.clamp(x, min=-5.0, max=5.0)
            y = torch.sqrt(cv.float() + self.epsilon) * y + cm.float()
        else:
            y = (x - cm.float()) / torch.sqrt(cv.float() + self.epsilon)
            y = torch.clamp(y, min=-5.0, max=5.0)
        return y

    def state(self):
        return {'running_mean': self.running_mean.clone(), 'running_var': self.running_var.clone(),
                'count': self.count.clone()}

    def load(self, st):
        self.running_mean = st['running_mean'].clone().double()
        self.running_var = st['running_var'].clone().double()
        self.count = st['count'].clone().long()


# ----------------------------------------------------------------------------------------------
# cross-rank running-stat merge -- rl_games/common/a2c_common.py:43-93
# ----------------------------------------------------------------------------------------------
class GeneralizedMovingStats:
    """EMA advantage normaliser, impl 'mean_std' only (algos_torch/moving_mean_std.py:7-150; created at a2c_common.py:473-475 with
    decay = adv_rms_momentum when normalize_rms_advantage).  State: int32 step (starts at 1), fp32 mean / mean of squares.
    Every update moves the statistics by the same decay whatever the number of valid rows; an all-invalid mask is a no-op."""

    def __init__(self, insize, decay=0.99, max=1e5, eps=0.0):
        self.decay, self.max, self.eps = decay, max, eps
        self.step = torch.ones(1, dtype=torch.int32)
        self.mean = torch.zeros(insize, dtype=torch.float32)
        self.sqrs = torch.zeros(insize, dtype=torch.float32)
        self.training = True

    def train(self):
        self.training = True

    def eval(self):
        self.training = False

    def get_mean_std(self):
        var = self.sqrs - self.mean.pow(2)
        return self.mean, torch.sqrt(torch.clamp_min(var, 1 / self.max ** 2) + self.eps)

    def __call__(self, x, mask=None, denorm=False):
        if self.training:
            xs = x
            if mask is not None:
                valid = mask.reshape(-1) > 0
                xs = x[valid] if bool(valid.any()) else None
            if xs is not None:
                m = self.decay
                self.step += 1
                self.mean.mul_(m).add_((1 - m) * torch.mean(xs, dim=0))
                self.sqrs.mul_(m).add_((1 - m) * torch.mean(xs * xs, dim=0))
        offset, invscale = self.get_mean_std()
        if denorm:
            return x.clone().mul_(invscale).add_(offset)
        return x.clone().sub_(offset).div_(invscale).clamp_(-5.0, 5.0)


def running_stats_totals(m):
    return (m.count.clone(), m.running_mean * m.count, (m.running_var + m.running_mean ** 2) * m.count)


def merge_rank_stats(m, all_reduce, snapshot=None):
    cur = running_stats_totals(m)
    if snapshot is None:
        deltas = [c.clone() for c in cur]
        base = [torch.zeros_like(c) for c in cur]
    else:
        deltas = [c - p for c, p in zip(cur, snapshot)]
        base = snapshot
    for t in deltas:
        all_reduce(t)
    n = base[0] + deltas[0]
    wm = base[1] + deltas[1]
    wsq = base[2] + deltas[2]
    m.count = n.clone()
    m.running_mean = wm / n
    m.running_var = (wsq / n - m.running_mean ** 2).clamp_(min=1e-8)
    return (n.clone(), wm.clone(), wsq.clone())


# ----------------------------------------------------------------------------------------------
# losses -- rl_games/common/common_losses.py:16-82, algos_torch/a2c_continuous.py:97-134,:241-257,
#           algos_torch/torch_ext.py:27-36
# ----------------------------------------------------------------------------------------------
def critic_loss(value_preds_batch, values, curr_e_clip, return_batch, clip_value):
    if clip_value:
        delta = values - value_preds_batch
        value_pred_clipped = value_preds_batch + delta.clamp(-curr_e_clip, curr_e_clip)
        value_losses = (values - return_batch) ** 2
        value_losses_clipped = (value_pred_clipped - return_batch) ** 2
        return torch.max(value_losses, value_losses_clipped)
    return (return_batch - values) ** 2


def smooth_clamp(x, mi, mx):
    return 1 / (1 + torch.exp((-(x - mi) / (mx - mi) + 0.5) * 4)) * (mx - mi) + mi


def actor_loss(old_neglogp, neglogp, advantage, is_ppo, curr_e_clip, smooth=False):
    if is_ppo:
        ratio = torch.exp(old_neglogp - neglogp)
        surr1 = advantage * ratio
        if smooth:
            surr2 = advantage * smooth_clamp(ratio, 1.0 - curr_e_clip, 1.0 + curr_e_clip)
        else:
            surr2 = advantage * torch.clamp(ratio, 1.0 - curr_e_clip, 1.0 + curr_e_clip)
        return torch.max(-surr1, -surr2)
    return neglogp * advantage


def bound_loss(mu, bounds_loss_coef):
    if bounds_loss_coef is not None:
        soft_bound = 1.1
        mu_loss_high = torch.clamp_min(mu - soft_bound, 0.0) ** 2
        mu_loss_low = torch.clamp_max(mu + soft_bound, 0.0) ** 2
        return (mu_loss_low + mu_loss_high).sum(axis=-1)
    return torch.zeros(mu.shape[0])


def reg_loss(mu, bounds_loss_coef):
    if bounds_loss_coef is not None:
        return (mu * mu).sum(axis=-1)
    return torch.zeros(mu.shape[0])


def policy_kl(p0_mu, p0_sigma, p1_mu, p1_sigma, reduce=True):
    c1 = torch.log(p1_sigma / p0_sigma + 1e-5)
    c2 = (p0_sigma ** 2 + (p1_mu - p0_mu) ** 2) / (2.0 * (p1_sigma ** 2 + 1e-5))
    c3 = -1.0 / 2.0
    kl = (c1 + c2 + c3).sum(dim=-1)
    return kl.mean() if reduce else kl


def neglogp_fn(x, mean, std, logstd):
    """rl_games/algos_torch/models.py:361-364."""
    return 0.5 * (((x - mean) / std) ** 2).sum(dim=-1) + 0.5 * math.log(2.0 * math.pi) * x.size(-1) \
        + logstd.sum(dim=-1)


def explained_variance(y_pred, y):
    """torch_ext.py:197-215 (unmasked)."""
    var_y = torch.var(y, unbiased=False)
    var_dy = torch.var(y - y_pred, unbiased=False)
    return 1.0 - var_dy / var_y


def policy_clip_fraction(new_neglogp, old_neglogp, clip_param):
    """torch_ext.py:217-227 (unmasked)."""
    logratio = old_neglogp - new_neglogp
    return torch.logical_or(logratio < math.log(1.0 - clip_param),
                            logratio > math.log(1.0 + clip_param)).float().mean()


# ----------------------------------------------------------------------------------------------
# LR schedulers -- rl_games/common/schedulers.py:19-58 (python floats == fp64)
# ----------------------------------------------------------------------------------------------
class AdaptiveScheduler:
    def __init__(self, kl_threshold=0.008, min_lr=1e-6, max_lr=1e-2, lr_multiplier=1.5):
        self.min_lr, self.max_lr = min_lr, max_lr
        self.kl_threshold, self.lr_multiplier = kl_threshold, lr_multiplier

    def update(self, current_lr, entropy_coef, epoch, frames, kl_dist):
        lr = current_lr
        if kl_dist > (2.0 * self.kl_threshold):
            lr = max(current_lr / self.lr_multiplier, self.min_lr)
        if kl_dist < (0.5 * self.kl_threshold):
            lr = min(current_lr * self.lr_multiplier, self.max_lr)
        return lr, entropy_coef


class LinearScheduler:
    def __init__(self, start_lr, min_lr=1e-6, max_steps=1000000, use_epochs=True, apply_to_entropy=False, start_entropy_coef=0.01,
                 min_entropy_coef=0.0001):
        self.start_lr, self.min_lr, self.max_steps, self.use_epochs = start_lr, min_lr, max_steps, use_epochs
        self.apply_to_entropy = apply_to_entropy            # schedulers.py:45-48, :57-58 (config key schedule_entropy)
        self.start_entropy_coef, self.min_entropy_coef = start_entropy_coef, min_entropy_coef

    def update(self, current_lr, entropy_coef, epoch, frames, kl_dist):
        steps = epoch if self.use_epochs else frames
        mul = max(0, self.max_steps - steps) / self.max_steps
        if self.apply_to_entropy:
            entropy_coef = self.min_entropy_coef + (self.start_entropy_coef - self.min_entropy_coef) * mul
        return self.min_lr + (self.start_lr - self.min_lr) * mul, entropy_coef


class IdentityScheduler:
    def update(self, current_lr, entropy_coef, epoch, frames, kl_dist):
        return current_lr, entropy_coef


# ----------------------------------------------------------------------------------------------
# model: A2CBuilder.Network (MLP trunk, shared) + ModelA2CContinuousLogStd
#   network_builder.py:219-348 (ctor, parameter order), :494-512 (forward), models.py:329-364
# parameter order (== reference model.parameters()):
#   sigma, actor_mlp.{0,2,4,..}.{weight,bias}, value.{weight,bias}, mu.{weight,bias}
# ----------------------------------------------------------------------------------------------
ACTIVATIONS = {'elu': F.elu, 'relu': F.relu, 'tanh': torch.tanh, 'None': lambda x: x}


RNN_NAMES = ['a2c_network.rnn.rnn.weight_ih_l0', 'a2c_network.rnn.rnn.weight_hh_l0', 'a2c_network.rnn.rnn.bias_ih_l0',
             'a2c_network.rnn.rnn.bias_hh_l0']


def param_names(n_layers, lstm=False, separate=False):
    """reference model.parameters() order: sigma, actor_mlp.*, [critic_mlp.* when `separate`], [rnn.rnn.*], value.*, mu.* (module
    registration order of A2CBuilder.Network.__init__, network_builder.py:219-325: actor_mlp placeholder is registered before self.rnn)"""
    names = ['a2c_network.sigma']
    for i in range(n_layers):
        names += [f'a2c_network.actor_mlp.{2 * i}.weight', f'a2c_network.actor_mlp.{2 * i}.bias']
    if separate:
        for i in range(n_layers):
            names += [f'a2c_network.critic_mlp.{2 * i}.weight', f'a2c_network.critic_mlp.{2 * i}.bias']
    if lstm:
        names += RNN_NAMES
    names += ['a2c_network.value.weight', 'a2c_network.value.bias', 'a2c_network.mu.weight', 'a2c_network.mu.bias']
    return names


def init_params(obs_dim, units, act_dim, value_size=1, seed=0, sigma_init=0.0, rnn_units=0, rnn_before_mlp=True):
    """Default torch.nn.Linear init (mlp initializer 'default' == nn.Identity on weights,
    biases zeroed: network_builder.py:332-340), mu_init default, sigma const 0."""
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}
    p['a2c_network.sigma'] = torch.full((act_dim,), float(sigma_init))
    ins = obs_dim
    if rnn_units and rnn_before_mlp:   # LSTM before the MLP (torch.nn.LSTM default init U(-1/sqrt(hid), 1/sqrt(hid)); mlp_init does not touch it)
        k = 1.0 / math.sqrt(rnn_units)
        p[RNN_NAMES[0]] = (torch.rand(4 * rnn_units, ins, generator=g) * 2 - 1) * k
        p[RNN_NAMES[1]] = (torch.rand(4 * rnn_units, rnn_units, generator=g) * 2 - 1) * k
        p[RNN_NAMES[2]] = (torch.rand(4 * rnn_units, generator=g) * 2 - 1) * k
        p[RNN_NAMES[3]] = (torch.rand(4 * rnn_units, generator=g) * 2 - 1) * k
        ins = rnn_units
    for i, u in enumerate(units):
        bound = 1.0 / math.sqrt(ins)
        p[f'a2c_network.actor_mlp.{2 * i}.weight'] = (torch.rand(u, ins, generator=g) * 2 - 1) * bound
        p[f'a2c_network.actor_mlp.{2 * i}.bias'] = torch.zeros(u)
        ins = u
    if rnn_units and not rnn_before_mlp:      # LSTM between the MLP and the heads
        k = 1.0 / math.sqrt(rnn_units)
        p[RNN_NAMES[0]] = (torch.rand(4 * rnn_units, ins, generator=g) * 2 - 1) * k
        p[RNN_NAMES[1]] = (torch.rand(4 * rnn_units, rnn_units, generator=g) * 2 - 1) * k
        p[RNN_NAMES[2]] = (torch.rand(4 * rnn_units, generator=g) * 2 - 1) * k
        p[RNN_NAMES[3]] = (torch.rand(4 * rnn_units, generator=g) * 2 - 1) * k
        ins = rnn_units
    bound = 1.0 / math.sqrt(ins)
    p['a2c_network.value.weight'] = (torch.rand(value_size, ins, generator=g) * 2 - 1) * bound
    p['a2c_network.value.bias'] = torch.zeros(value_size)
    p['a2c_network.mu.weight'] = (torch.rand(act_dim, ins, generator=g) * 2 - 1) * bound
    p['a2c_network.mu.bias'] = torch.zeros(act_dim)
    return p


def network_forward(p, obs, n_layers, activation='elu', matmul_dtype=None, heads_after=None):
    """a2c network forward (non-rnn, non-separate): network_builder.py:448-512.
    matmul_dtype=torch.bfloat16 emulates bf16 autocast of nn.Linear (a2c_continuous.py:173)."""
    act = ACTIVATIONS[activation]

    def lin(x, w, b):
        if matmul_dtype is not None:
            return F.linear(x.to(matmul_dtype), w.to(matmul_dtype), b.to(matmul_dtype))
        return F.linear(x, w, b)
    out = obs
    for i in range(n_layers):
        out = act(lin(out, p[f'a2c_network.actor_mlp.{2 * i}.weight'], p[f'a2c_network.actor_mlp.{2 * i}.bias']))
    if heads_after is not None:          # rnn after the MLP (before_mlp: False, network_builder.py:466-492): trunk -> rnn -> heads
        out = heads_after(out)
    c_out = out
    if 'a2c_network.critic_mlp.0.weight' in p:      # separate: True (network_builder.py:494-512): the value head reads a trunk of its own
        c_out = obs
        for i in range(n_layers):
            c_out = act(lin(c_out, p[f'a2c_network.critic_mlp.{2 * i}.weight'], p[f'a2c_network.critic_mlp.{2 * i}.bias']))
    value = lin(c_out, p['a2c_network.value.weight'], p['a2c_network.value.bias'])
    mu = lin(out, p['a2c_network.mu.weight'], p['a2c_network.mu.bias'])
    logstd = mu * 0 + p['a2c_network.sigma']
    return mu, logstd, value


def lstm_with_dones(p, x, h, c, dones):
    """LSTMWithDones (common/layers/recurrent.py:20-80): single-layer torch.nn.LSTM cell maths, states multiplied by
    (1 - done[t]) before step t (RnnWithDones.forward :26-58 is equivalent to masking at every step: inside a done-free
    segment the mask is all ones).  x [T,B,F], h/c [B,hid], dones [T,B] or None.  Gate order i,f,g,o."""
    w_ih, w_hh, b_ih, b_hh = (p[n] for n in RNN_NAMES)
    outs = []
    for t in range(x.shape[0]):
        if dones is not None:
            nd = (1.0 - dones[t].float()).unsqueeze(1)
            h, c = h * nd, c * nd
        gates = F.linear(x[t], w_ih, b_ih) + F.linear(h, w_hh, b_hh)
        i, f, g, o = gates.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs), h, c


class OracleModel:
    """ModelA2CContinuousLogStd.Network (models.py:304-364) + BaseModelNetwork (:38-63)."""

    def __init__(self, params, obs_dim, units, act_dim, normalize_input=True, normalize_value=True,
                 activation='elu', value_size=1, matmul_dtype=None, rnn_units=0, rnn_before_mlp=True, min_sigma=0.0):
        self.p = {k: v.clone().float().requires_grad_(True) for k, v in params.items()}
        self.min_sigma = float(min_sigma)      # 'exp' sigma parametrisation with a floor (models.py:272-300, network_builder.py:312)
        self.rnn_units = rnn_units
        self.rnn_before_mlp = rnn_before_mlp
        self.names = param_names(len(units), lstm=rnn_units > 0, separate='a2c_network.critic_mlp.0.weight' in params)
        self.n_layers = len(units)
        self.activation = activation
        self.normalize_input, self.normalize_value = normalize_input, normalize_value
        self.running_mean_std = RunningMeanStd((obs_dim,)) if normalize_input else None
        self.value_mean_std = RunningMeanStd((value_size,)) if normalize_value else None
        self.matmul_dtype = matmul_dtype

    def parameters(self):
        return [self.p[n] for n in self.names]

    def norm_obs(self, obs):
        with torch.no_grad():
            return self.running_mean_std(obs) if self.normalize_input else obs

    def denorm_value(self, value):
        with torch.no_grad():
            return self.value_mean_std(value, denorm=True) if self.normalize_value else value

    def forward(self, obs, is_train, prev_actions=None, noise=None, rnn_states=None, seq_length=1, dones=None):
        """rnn (before_mlp) branch: network_builder.py:452-492 -- [B,F] -> [num_seqs, seq, F] -> transpose -> LSTM -> back."""
        obs = self.norm_obs(obs)
        new_states = None

        def rnn(x_flat):
            """[B,F] -> [num_seqs, seq, F] -> transpose -> LSTM with dones -> back to [B, hid]"""
            nonlocal new_states
            B = x_flat.shape[0]
            num_seqs = B // seq_length
            x = x_flat.reshape(num_seqs, seq_length, -1).transpose(0, 1)
            d = None if dones is None else dones.reshape(num_seqs, seq_length).transpose(0, 1)
            out, h, c = lstm_with_dones(self.p, x, rnn_states[0][0], rnn_states[1][0], d)
            new_states = (h.unsqueeze(0), c.unsqueeze(0))
            return out.transpose(0, 1).contiguous().reshape(B, -1)
        after = None
        if self.rnn_units and self.rnn_before_mlp:
            obs = rnn(obs)
        elif self.rnn_units:
            after = rnn
        mu, logstd, value = network_forward(self.p, obs, self.n_layers, self.activation, self.matmul_dtype, heads_after=after)
        self.last_rnn_states = new_states
        mu, logstd, value = mu.float(), logstd.float(), value.float()
        sigma = torch.exp(logstd)
        if self.min_sigma > 0:                 # models.py:296-300: the floor is ADDED, and the log-std is recomputed from the final sigma
            sigma = sigma + self.min_sigma
            logstd = torch.log(sigma)
        if is_train:
            entropy = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(sigma)).sum(dim=-1)
            prev_neglogp = neglogp_fn(prev_actions, mu, sigma, logstd)
            return {'prev_neglogp': prev_neglogp, 'values': value, 'entropy': entropy, 'mus': mu, 'sigmas': sigma}
        # Normal(mu, sigma).sample() == mu + sigma * eps ; eps supplied so both sides share it
        actions = mu + sigma * noise
        neglogp = neglogp_fn(actions, mu, sigma, logstd)
        return {'neglogpacs': neglogp, 'values': self.denorm_value(value), 'actions': actions,
                'mus': mu, 'sigmas': sigma}


# ----------------------------------------------------------------------------------------------
# Adam (torch.optim.Adam(eps=1e-8, weight_decay, fused=True) semantics, a2c_continuous.py:44-48)
# and clip_grad_norm_ (a2c_common.py:511-512)
# ----------------------------------------------------------------------------------------------
class Adam:
    def __init__(self, params, lr, eps=1e-8, weight_decay=0.0, betas=(0.9, 0.999)):
        self.params = params
        self.lr, self.eps, self.wd, self.betas = lr, eps, weight_decay, betas
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.step_count = 0

    @torch.no_grad()
    def step(self, grads):
        self.step_count += 1
        b1, b2 = self.betas
        bc1 = 1 - b1 ** self.step_count
        bc2 = 1 - b2 ** self.step_count
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            if self.wd != 0:
                g = g + self.wd * p
            m.lerp_(g, 1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            step_size = self.lr / bc1
            denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(m, denom, value=-step_size)


def clip_grad_norm(grads, max_norm):
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
    clip_coef = max_norm / (total + 1e-6)
    clip_coef_clamped = torch.clamp(clip_coef, max=1.0)
    return [g * clip_coef_clamped for g in grads], total


# ----------------------------------------------------------------------------------------------
# reward shaper -- rl_games/common/tr_helpers.py:16-42
# ----------------------------------------------------------------------------------------------
def shape_rewards(r, scale_value=1.0, shift_value=0.0, min_val=-float('inf'), max_val=float('inf'), log_val=False):
    r = (r + shift_value) * scale_value
    r = torch.clamp(r, min_val, max_val)
    return torch.log(r) if log_val else r


# ----------------------------------------------------------------------------------------------
# AverageMeter -- rl_games/algos_torch/torch_ext.py:326-352
# ----------------------------------------------------------------------------------------------
class AverageMeter:
    def __init__(self, in_shape, max_size):
        self.max_size, self.current_size = max_size, 0
        self.mean = torch.zeros(in_shape, dtype=torch.float32)

    def update(self, values):
        size = values.size()[0]
        if size == 0:
            return
        new_mean = torch.mean(values.float(), dim=0)
        size = min(max(size, 0), self.max_size)
        old_size = min(self.max_size - size, self.current_size)
        size_sum = old_size + size
        self.current_size = size_sum
        self.mean = (self.mean * old_size + new_mean * size) / size_sum


# ----------------------------------------------------------------------------------------------
# The agent: A2CBase / ContinuousA2CBase / A2CAgent restated.
#   play_steps            a2c_common.py:985-1069
#   prepare_dataset       a2c_common.py:1586-1660
#   train_epoch           a2c_common.py:1517-1584
#   calc_gradients        a2c_continuous.py:136-234
#   trancate_gradients_and_step  a2c_common.py:493-514
# ----------------------------------------------------------------------------------------------
DEFAULT_CFG = dict(
    gamma=0.99, tau=0.95, e_clip=0.2, clip_value=True, critic_coef=2.0, entropy_coef=0.0,
    bounds_loss_coef=0.0, bound_loss_type='regularisation', use_smooth_clamp=True,
    truncate_grads=True, grad_norm=1.0, learning_rate=3e-4, lr_schedule='adaptive', kl_threshold=0.008,
    min_lr=1e-6, max_lr=1e-2, lr_multiplier=1.5, schedule_type='per_minibatch',
    normalize_input=True, normalize_value=True, normalize_advantage=True, value_bootstrap=True,
    mini_epochs=4, weight_decay=0.0, ppo=True, reward_scale=1.0, reward_shift=0.0, reward_min=-float('inf'), reward_max=float('inf'),
    reward_log=False, max_epochs=-1, max_frames=-1, schedule_entropy=False, actions_low=-1.0, actions_high=1.0, seq_length=4, rnn_units=0, zero_rnn_on_done=True,
    games_to_track=100, activation='elu', clip_actions=True, mask_autoreset_rows=False,
    normalize_rms_advantage=False, adv_rms_momentum=0.5, min_sigma=0.0,
)


# ----------------------------------------------------------------------------------------------
# Central value (asymmetric critic) -- algos_torch/central_value.py:14-383, models.py:425-464 (ModelCentralValue),
# A2CBuilder.Network with central_value: True (network_builder.py:498-499: returns value only).  SURVEY 8f rank 1.
# ----------------------------------------------------------------------------------------------
def cv_param_names(n_layers):
    names = []
    for i in range(n_layers):
        names += [f'a2c_network.actor_mlp.{2 * i}.weight', f'a2c_network.actor_mlp.{2 * i}.bias']
    return names + ['a2c_network.value.weight', 'a2c_network.value.bias']


class CentralValueOracle:
    """CentralValueTrain: its own MLP on `states`, own obs normaliser, the (shared) value normaliser, own Adam and lr."""

    def __init__(self, params, state_dim, units, cv_cfg, normalize_value, num_actors, horizon, activation='elu'):
        self.p = {k: v.clone().float().requires_grad_(True) for k, v in params.items()}
        self.names = cv_param_names(len(units))
        self.n_layers, self.activation = len(units), activation
        c = self.cfg = dict(cv_cfg)
        self.normalize_input, self.normalize_value = c['normalize_input'], normalize_value
        self.running_mean_std = RunningMeanStd((state_dim,)) if self.normalize_input else None
        self.value_mean_std = RunningMeanStd((1,)) if normalize_value else None
        self.lr = float(c['learning_rate'])
        self.mini_epoch = c['mini_epochs']
        self.batch_size = num_actors * horizon
        self.minibatch_size = c['minibatch_size']
        self.num_minibatches = self.batch_size // self.minibatch_size
        self.clip_value, self.e_clip = c['clip_value'], c.get('e_clip', 0.2)
        self.truncate_grads, self.grad_norm = c.get('truncate_grads', False), c.get('grad_norm', 1)
        self.optimizer = Adam([self.p[n] for n in self.names], self.lr, eps=1e-8, weight_decay=c.get('weight_decay', 0.0))
        self.scheduler = IdentityScheduler()          # lr_schedule 'linear' is the only other option (central_value.py:55-62)
        self.epoch_num, self.frame = 0, 0
        self.dataset = None
        self.all_reduce, self.world_size = None, 1     # multi-GPU: set both (central_value.py:322-337)

    def parameters(self):
        return [self.p[n] for n in self.names]

    def forward(self, states, is_train):
        x = states
        if self.normalize_input:
            with torch.no_grad():
                x = self.running_mean_std(x)
        act = ACTIVATIONS[self.activation]
        for i in range(self.n_layers):
            x = act(F.linear(x, self.p[f'a2c_network.actor_mlp.{2 * i}.weight'], self.p[f'a2c_network.actor_mlp.{2 * i}.bias']))
        value = F.linear(x, self.p['a2c_network.value.weight'], self.p['a2c_network.value.bias'])
        if not is_train and self.normalize_value:
            with torch.no_grad():
                value = self.value_mean_std(value, denorm=True)
        return value

    @torch.no_grad()
    def get_value(self, states):               # central_value.py:207-229: self.eval() first
        if self.normalize_input:
            self.running_mean_std.eval()
        if self.normalize_value:
            self.value_mean_std.eval()
        return self.forward(states, is_train=False)

    def update_dataset(self, d):               # central_value.py:153-176
        self.dataset = d

    def train_net(self):                       # central_value.py:246-274
        loss = 0.0
        self.last_losses = []
        for _ in range(self.mini_epoch):
            if self.cfg.get('freeze_critic', False):
                break
            for i in range(self.num_minibatches):
                # train_critic -> self.train(): nn.Module.train re-arms BOTH normalisers of the critic model before every
                # minibatch; the value normaliser in train mode would update from the critic's own outputs in forward... it is
                # only ever called through denorm in eval mode, so only the obs normaliser moves here
                if self.normalize_input:
                    self.running_mean_std.train()
                s, e = i * self.minibatch_size, (i + 1) * self.minibatch_size
                values = self.forward(self.dataset['obs'][s:e], is_train=True)
                masks = self.dataset.get('rnn_masks')
                l = critic_loss(self.dataset['old_values'][s:e], values, self.e_clip, self.dataset['returns'][s:e], self.clip_value)
                losses, _ = apply_masks([l], None if masks is None else masks[s:e])
                l = losses[0]
                params = self.parameters()
                for p in params:
                    p.grad = None
                l.backward()
                grads = [p.grad for p in params]
                if self.all_reduce is not None:         # central_value.py:322-337: cat -> all_reduce(SUM) -> / world -> scatter
                    flat = torch.cat([g.reshape(-1) for g in grads])
                    self.all_reduce(flat)
                    off, new = 0, []
                    for g in grads:
                        new.append(flat[off:off + g.numel()].view_as(g) / self.world_size)
                        off += g.numel()
                    grads = new
                if self.truncate_grads:
                    grads, _ = clip_grad_norm(grads, self.grad_norm)
                self.optimizer.lr = self.lr
                self.optimizer.step(grads)
                self.last_losses.append(l.detach())
                loss += float(l.detach())
            if self.normalize_input:
                self.running_mean_std.eval()
        self.epoch_num += 1
        self.lr, _ = self.scheduler.update(self.lr, 0, self.epoch_num, self.frame, 0)
        self.frame += self.batch_size
        return loss / (self.mini_epoch * self.num_minibatches)


class OracleAgent:
    def __init__(self, env, params, obs_dim, act_dim, units, num_actors, horizon, minibatch_size,
                 cfg: Optional[dict] = None, matmul_dtype=None, all_reduce=None, world_size=1, central_value=None):
        self.cv = central_value            # CentralValueOracle or None (a2c_common.py:250-251 has_central_value)
        self.cfg = dict(DEFAULT_CFG)
        self.cfg.update(cfg or {})
        c = self.cfg
        self.env = env
        self.N, self.H, self.D, self.A = num_actors, horizon, obs_dim, act_dim
        self.batch_size = self.N * self.H
        self.minibatch_size = minibatch_size
        assert self.batch_size % minibatch_size == 0
        self.num_minibatches = self.batch_size // minibatch_size
        self.model = OracleModel(params, obs_dim, units, act_dim, c['normalize_input'], c['normalize_value'],
                                 c['activation'], matmul_dtype=matmul_dtype, rnn_units=c['rnn_units'],
                                 rnn_before_mlp=c.get('rnn_before_mlp', True), min_sigma=c.get('min_sigma', 0.0))
        self.is_rnn = c['rnn_units'] > 0
        self.seq_length = c['seq_length']
        self.last_lr = float(c['learning_rate'])
        self.entropy_coef = c['entropy_coef']
        self.optimizer = Adam(self.model.parameters(), self.last_lr, eps=1e-8, weight_decay=c['weight_decay'])
        if c['lr_schedule'] == 'adaptive':
            self.scheduler = AdaptiveScheduler(c['kl_threshold'], c['min_lr'], c['max_lr'], c['lr_multiplier'])
        elif c['lr_schedule'] == 'linear' and (c['max_epochs'] != -1 or c['max_frames'] != -1):       # a2c_common.py:314-332
            use_epochs = c['max_epochs'] != -1
            self.scheduler = LinearScheduler(float(c['learning_rate']), c['min_lr'], c['max_epochs'] if use_epochs else c['max_frames'],
                                             use_epochs, c['schedule_entropy'], c['entropy_coef'])
        else:
            self.scheduler = IdentityScheduler()
        self.all_reduce, self.world_size = all_reduce, world_size
        self.epoch_num, self.frame = 0, 0
        self.game_rewards = AverageMeter(1, c['games_to_track'])
        self.game_shaped_rewards = AverageMeter(1, c['games_to_track'])
        self.game_lengths = AverageMeter(1, c['games_to_track'])
        self.mask_autoreset_rows = c['mask_autoreset_rows']
        self._autoreset_prev_dones = None
        self.actions_low = torch.full((act_dim,), float(c['actions_low']))        # the env's action_space bounds (a2c_common.py:1496-1497)
        self.actions_high = torch.full((act_dim,), float(c['actions_high']))
        self.init_tensors()

    # a2c_common.py:634-670
    def init_tensors(self):
        H, N, D, A = self.H, self.N, self.D, self.A
        self.buf = {
            'obses': torch.zeros(H, N, D), 'rewards': torch.zeros(H, N, 1), 'values': torch.zeros(H, N, 1),
            'neglogpacs': torch.zeros(H, N), 'dones': torch.zeros(H, N, dtype=torch.uint8),
            'actions': torch.zeros(H, N, A), 'mus': torch.zeros(H, N, A), 'sigmas': torch.zeros(H, N, A),
        }
        self.current_rewards = torch.zeros(N, 1)
        self.current_shaped_rewards = torch.zeros(N, 1)
        self.current_lengths = torch.zeros(N)
        self.dones = torch.ones(N, dtype=torch.uint8)
        self.obs = None
        if self.is_rnn:   # a2c_common.py:652-660
            hid = self.cfg['rnn_units']
            self.rnn_states = [torch.zeros(1, N, hid), torch.zeros(1, N, hid)]
            self.mb_rnn_states = [torch.zeros(H // self.seq_length, 1, N, hid) for _ in range(2)]

    def _take_obs(self, o):
        """envs with a central value return {'obs': ..., 'states': ...} (a2c_common.py:1010-1011)"""
        if isinstance(o, dict):
            self.states = o['states']
            return o['obs']
        return o

    def env_reset(self):
        self._autoreset_prev_dones = None
        return self._take_obs(self.env.reset())

    # a2c_common.py:1500-1510 + :144-148
    def preprocess_actions(self, actions):
        if self.cfg['clip_actions']:
            clamped = torch.clamp(actions, -1.0, 1.0)
            d = (self.actions_high - self.actions_low) / 2.0
            m = (self.actions_high + self.actions_low) / 2.0
            return clamped * d + m
        return actions

    @torch.no_grad()
    def play_steps(self, noise):
        """noise: [H, N, A] standard-normal draws used for action sampling."""
        c = self.cfg
        H, N = self.H, self.N
        if self.mask_autoreset_rows:
            mb_valid = torch.ones(H, N)
        for n in range(H):
            if self.is_rnn and n % self.seq_length == 0:       # play_steps_rnn: a2c_common.py:1081-1083
                for s_, mb_s in zip(self.rnn_states, self.mb_rnn_states):
                    mb_s[n // self.seq_length] = s_
            res = self.model.forward(self.obs, is_train=False, noise=noise[n], rnn_states=self.rnn_states if self.is_rnn else None)
            if self.cv is not None:                 # get_action_values: a2c_common.py:593-600
                res['values'] = self.cv.get_value(self.states)
                self.buf.setdefault('states', torch.zeros(H, N, self.states.shape[-1]))[n] = self.states
            if self.is_rnn:
                self.rnn_states = [t.clone() for t in self.model.last_rnn_states]
            self.buf['obses'][n] = self.obs
            self.buf['dones'][n] = self.dones
            if self.mask_autoreset_rows:
                prev = self._autoreset_prev_dones
                if prev is None:
                    prev = torch.zeros_like(self.dones)
                mb_valid[n] = 1.0 - prev.float()
                if self.is_rnn and c['zero_rnn_on_done']:
                    # a2c_common.py:1097-1106: a filler reset row's forward absorbed the dead episode's terminal obs into the freshly
                    # zeroed state -- re-zero, so that the first real row of the new episode starts clean
                    reset_idx = prev.nonzero(as_tuple=False)
                    if len(reset_idx) > 0:
                        for s_ in self.rnn_states:
                            s_[:, reset_idx, :] = 0
            for k in ['actions', 'neglogpacs', 'values', 'mus', 'sigmas']:
                self.buf[k][n] = res[k]
            o, rewards, self.dones, infos = self.env.step(self.preprocess_actions(res['actions']))
            self.obs = self._take_obs(o)
            rewards = rewards.unsqueeze(1)
            if self.mask_autoreset_rows:
                self._autoreset_prev_dones = self.dones.clone()
            shaped = shape_rewards(rewards, c['reward_scale'], c['reward_shift'], c['reward_min'], c['reward_max'], c['reward_log'])
            if c['value_bootstrap'] and 'time_outs' in infos:
                shaped = shaped + c['gamma'] * res['values'] * infos['time_outs'].unsqueeze(1).float()
            self.buf['rewards'][n] = shaped
            if self.mask_autoreset_rows:
                live = mb_valid[n]
                self.current_rewards.add_(rewards * live.unsqueeze(1))
                self.current_shaped_rewards.add_(shaped * live.unsqueeze(1))
                self.current_lengths.add_(live)
            else:
                self.current_rewards.add_(rewards)
                self.current_shaped_rewards.add_(shaped)
                self.current_lengths.add_(1)
            done_idx = self.dones.nonzero(as_tuple=False)
            if self.is_rnn and len(done_idx) > 0 and c['zero_rnn_on_done']:     # a2c_common.py:1150-1153
                for s_ in self.rnn_states:
                    s_[:, done_idx, :] = 0
            self.game_rewards.update(self.current_rewards[done_idx])
            self.game_shaped_rewards.update(self.current_shaped_rewards[done_idx])
            self.game_lengths.update(self.current_lengths[done_idx])
            nd = (1.0 - self.dones.float()).unsqueeze(1)
            self.current_rewards.mul_(nd)
            self.current_shaped_rewards.mul_(nd)
            self.current_lengths.mul_(nd.squeeze(1))
        if self.cv is not None:                     # get_values: a2c_common.py:603-615
            last_values = self.cv.get_value(self.states)
        else:
            last_values = self.model.forward(self.obs, is_train=False, noise=torch.zeros(N, self.A),
                                             rnn_states=self.rnn_states if self.is_rnn else None)['values']
        fdones = self.dones.float()
        mb_fdones = self.buf['dones'].float()
        mb_advs = gae(self.buf['rewards'], self.buf['values'], mb_fdones, last_values, fdones, c['gamma'], c['tau'])
        mb_returns = mb_advs + self.buf['values']
        batch = {k: swap_and_flatten01(self.buf[k]) for k in
                 ['actions', 'neglogpacs', 'values', 'mus', 'sigmas', 'obses', 'dones']}
        batch['returns'] = swap_and_flatten01(mb_returns)
        batch['mb_advs'] = mb_advs
        if self.cv is not None:
            batch['states'] = swap_and_flatten01(self.buf['states'])
        if self.mask_autoreset_rows:
            batch['rnn_masks'] = swap_and_flatten01(mb_valid)
            if self.is_rnn and c['zero_rnn_on_done']:
                # a2c_common.py:1180-1191: the batch's dones feed only the train-time state reset; also reset ENTERING the first real row
                # after a filler reset row (GAE above used the buffer's own dones)
                rnn_dones = self.buf['dones'].clone()
                garbage = (mb_valid == 0.0)
                rnn_dones[1:] = torch.maximum(rnn_dones[1:], garbage[:-1].to(rnn_dones.dtype))
                batch['dones'] = swap_and_flatten01(rnn_dones)
        if self.is_rnn:    # a2c_common.py:1193-1199
            states = []
            for mb_s in self.mb_rnn_states:
                t_size = mb_s.size()[0] * mb_s.size()[2]
                states.append(mb_s.permute(1, 2, 0, 3).reshape(-1, t_size, mb_s.size()[3]))
            batch['rnn_states'] = states
        return batch

    def prepare_dataset(self, batch):
        c = self.cfg
        returns, values = batch['returns'], batch['values']
        rnn_masks = batch.get('rnn_masks', None)
        advantages = returns - values
        if c['normalize_value']:
            # with a central value the agent's value normaliser IS the critic model's (a2c_continuous.py:72-73)
            cv = getattr(self, 'cv', None)          # (tests build bare agents with __new__ to call this method alone)
            vms = cv.value_mean_std if cv is not None else self.model.value_mean_std
            if rnn_masks is not None:
                valid = rnn_masks.bool()
                vms.train()
                vms(values[valid])
                vms(returns[valid])
                vms.eval()
                values = vms(values)
                returns = vms(returns)
            else:
                vms.train()
                values = vms(values)
                returns = vms(returns)
                vms.eval()
        advantages = torch.sum(advantages, axis=1)
        if c['normalize_advantage']:      # a2c_common.py:1622-1632
            if c.get('normalize_rms_advantage', False):
                if getattr(self, 'advantage_mean_std', None) is None:
                    self.advantage_mean_std = GeneralizedMovingStats((1,), decay=c.get('adv_rms_momentum', 0.5))
                advantages = self.advantage_mean_std(advantages, mask=rnn_masks) if rnn_masks is not None else self.advantage_mean_std(advantages)
            elif rnn_masks is not None:
                advantages = normalization_with_masks(advantages, rnn_masks)
            else:
                advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
        self.dataset = {
            'old_values': values, 'old_logp_actions': batch['neglogpacs'], 'advantages': advantages,
            'returns': returns, 'actions': batch['actions'], 'obs': batch['obses'], 'dones': batch['dones'],
            'rnn_masks': rnn_masks, 'mu': batch['mus'].clone(), 'sigma': batch['sigmas'].clone(),
            'rnn_states': batch.get('rnn_states', None),
        }
        if getattr(self, 'cv', None) is not None:   # a2c_common.py:1651-1660
            self.cv.update_dataset({'old_values': values, 'advantages': advantages, 'returns': returns, 'actions': batch['actions'],
                                    'obs': batch['states'], 'dones': batch['dones'], 'rnn_masks': rnn_masks})

    def get_minibatch(self, i):
        s, e = i * self.minibatch_size, (i + 1) * self.minibatch_size
        self.last_range = (s, e)
        mb = {k: (v[s:e] if v is not None else None) for k, v in self.dataset.items() if k != 'rnn_states'}
        if self.is_rnn:    # datasets.py:63-73
            ng = self.minibatch_size // self.seq_length
            mb['rnn_states'] = [st[:, i * ng:(i + 1) * ng, :].contiguous() for st in self.dataset['rnn_states']]
        return mb

    def calc_losses(self, mb, res):
        c = self.cfg
        rnn_masks = mb.get('rnn_masks', None)
        a_loss = actor_loss(mb['old_logp_actions'], res['prev_neglogp'], mb['advantages'], c['ppo'], c['e_clip'],
                            smooth=c['use_smooth_clamp'])
        c_loss = critic_loss(mb['old_values'], res['values'], c['e_clip'], mb['returns'], c['clip_value'])
        if c['bound_loss_type'] == 'regularisation':
            b_loss = reg_loss(res['mus'], c['bounds_loss_coef'])
        elif c['bound_loss_type'] == 'bound':
            b_loss = bound_loss(res['mus'], c['bounds_loss_coef'])
        else:
            b_loss = torch.zeros(1)
        losses, sum_mask = apply_masks([a_loss.unsqueeze(1), c_loss, res['entropy'].unsqueeze(1),
                                        b_loss.unsqueeze(1)], rnn_masks)
        a_loss, c_loss, entropy, b_loss = losses
        bounds_coef = c['bounds_loss_coef'] if c['bounds_loss_coef'] is not None else 0.0
        loss = a_loss + 0.5 * c_loss * c['critic_coef'] - entropy * self.entropy_coef + b_loss * bounds_coef
        return loss, a_loss, c_loss, entropy, b_loss

    def train_actor_critic(self, mb):
        c = self.cfg
        rnn_masks = mb.get('rnn_masks', None)
        if self.is_rnn:
            res = self.model.forward(mb['obs'], is_train=True, prev_actions=mb['actions'], rnn_states=mb['rnn_states'],
                                     seq_length=self.seq_length, dones=mb['dones'] if c['zero_rnn_on_done'] else None)
        else:
            res = self.model.forward(mb['obs'], is_train=True, prev_actions=mb['actions'])
        loss, a_loss, c_loss, entropy, b_loss = self.calc_losses(mb, res)
        params = self.model.parameters()
        for p in params:
            p.grad = None
        loss.backward()
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
        # trancate_gradients_and_step
        if self.all_reduce is not None:
            flat = torch.cat([g.view(-1) for g in grads])
            self.all_reduce(flat)
            off = 0
            new = []
            for g in grads:
                new.append(flat[off:off + g.numel()].view_as(g) / self.world_size)
                off += g.numel()
            grads = new
        self.last_grad_norm = None
        if c['truncate_grads']:
            grads, self.last_grad_norm = clip_grad_norm(grads, c['grad_norm'])
        self.last_grads = [g.clone() for g in grads]
        self.optimizer.lr = self.last_lr
        self.optimizer.step(grads)
        with torch.no_grad():
            mu, sigma = res['mus'].detach(), res['sigmas'].detach()
            kl = policy_kl(mu, sigma, mb['mu'], mb['sigma'], reduce=rnn_masks is None)
            if rnn_masks is not None:
                kl = (kl * rnn_masks).sum() / rnn_masks.sum().clamp(min=1.0)
        return a_loss.detach(), c_loss.detach(), entropy.detach(), kl, mu, sigma, b_loss.detach()

    def train_epoch(self, noise):
        """One full epoch.  Returns dict of per-minibatch stats (lists)."""
        self.epoch_num += 1
        batch = self.play_steps(noise)
        self.prepare_dataset(batch)
        out = {'a_loss': [], 'c_loss': [], 'entropy': [], 'kl': [], 'b_loss': [], 'lr': [], 'batch': batch}
        if self.cv is not None:                     # a2c_common.py:1536-1537 train_central_value() before the actor's mini-epochs
            out['cv_loss'] = self.cv.train_net()
        for mini_ep in range(self.cfg['mini_epochs']):
            ep_kls = []
            for i in range(self.num_minibatches):
                # train_actor_critic -> set_train() -> model.train() (a2c_continuous.py:236-239,
                # a2c_common.py:560-563) puts running_mean_std back in TRAIN mode before EVERY
                # minibatch, undoing the `.eval()` at a2c_common.py:1575-1576: the reference updates the
                # obs statistics on every minibatch of every mini-epoch (golden count = 1 + B*mini_epochs).
                if self.model.normalize_input:
                    self.model.running_mean_std.train()
                mb = self.get_minibatch(i)
                a, cl, ent, kl, mu, sigma, b = self.train_actor_critic(mb)
                s, e = self.last_range
                self.dataset['mu'][s:e] = mu
                self.dataset['sigma'][s:e] = sigma
                out['a_loss'].append(a); out['c_loss'].append(cl); out['entropy'].append(ent)
                out['kl'].append(kl); out['b_loss'].append(b); out['lr'].append(self.last_lr)
                ep_kls.append(kl)
                if self.cfg['schedule_type'] == 'per_minibatch':
                    av_kl = kl
                    if self.all_reduce is not None:
                        av_kl = kl.clone()
                        self.all_reduce(av_kl)
                        av_kl = av_kl / self.world_size
                    self.last_lr, self.entropy_coef = self.scheduler.update(
                        self.last_lr, self.entropy_coef, self.epoch_num, self.frame, av_kl.item())
            if self.cfg['schedule_type'] == 'standard':
                # a2c_common.py:1565-1571: ONE scheduler step per mini-epoch on the mean of its minibatch KLs (torch_ext.mean_list)
                av_kl = torch.stack(ep_kls).mean()
                if self.all_reduce is not None:
                    self.all_reduce(av_kl)
                    av_kl = av_kl / self.world_size
                self.last_lr, self.entropy_coef = self.scheduler.update(
                    self.last_lr, self.entropy_coef, self.epoch_num, self.frame, av_kl.item())
            if self.model.normalize_input:
                self.model.running_mean_std.eval()
        self.frame += self.batch_size * self.world_size
        return out


# ----------------------------------------------------------------------------------------------
# Synthetic measurement env: obs ~ N(0,1), reward = -||a||^2, done at t>=100 or
# Bernoulli(0.01); time_outs = (t >= max_len) subset.  CPU/torch edition, used by the oracle arm.
# ----------------------------------------------------------------------------------------------
class SyntheticEnvCPU:
    def __init__(self, num_envs, obs_dim, act_dim, seed=5, max_len=100, p_done=0.01):
        self.N, self.D, self.A = num_envs, obs_dim, act_dim
        self.g = torch.Generator().manual_seed(seed)
        self.max_len, self.p_done = max_len, p_done
        self.t = torch.zeros(num_envs, dtype=torch.int32)

    def reset(self):
        self.t.zero_()
        return torch.randn(self.N, self.D, generator=self.g)

    def step(self, actions):
        rew = -(actions * actions).sum(-1)
        self.t += 1
        time_out = self.t >= self.max_len
        term = torch.rand(self.N, generator=self.g) < self.p_done
        done = time_out | term
        self.t[done] = 0
        obs = torch.randn(self.N, self.D, generator=self.g)
        return obs, rew, done.to(torch.uint8), {'time_outs': time_out & ~term}


class TapeEnv:
    """Deterministic env replaying pre-generated obs/dones/timeouts; reward = -||a||^2 * 0.1.
    Used so the oracle and the CUDA agent see identical trajectories in parity tests."""

    def __init__(self, obs_tape, done_tape, timeout_tape):
        self.obs_tape, self.done_tape, self.timeout_tape = obs_tape, done_tape, timeout_tape
        self.i = 0

    def reset(self):
        self.i = 0
        return self.obs_tape[0].clone()

    def step(self, actions):
        rew = -(actions * actions).sum(-1) * 0.1
        self.i += 1
        j = self.i % self.obs_tape.shape[0]
        return (self.obs_tape[j].clone(), rew, self.done_tape[j].clone(),
                {'time_outs': self.timeout_tape[j].clone()})


def make_tapes(T, N, D, seed=0, p_done=0.05, p_timeout=0.02):
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(T, N, D, generator=g) * 1.5 + 0.3
    done = (torch.rand(T, N, generator=g) < p_done)
    tout = (torch.rand(T, N, generator=g) < p_timeout) & done
    return obs, done.to(torch.uint8), tout
