"""TEST INFRASTRUCTURE ONLY - CPU restatement (oracle) of the learner half of the hot path.

Restates rsl_rl/rsl_rl/modules/actor_critic_decoder.py:91-302 (Vae), :305-551 (ActorCriticDecoder),
rsl_rl/rsl_rl/storage/rollout_storage.py:36-214 (RolloutStorage), rsl_rl/rsl_rl/algorithms/ppo.py:42-357
(PPO) and rsl_rl/rsl_rl/runners/on_policy_runner.py:86-163 (the learn loop) in plain torch on CPU,
with autograd as the differentiation oracle for the hand-written CUDA backward.  All randomness is
drawn through an `oracle.rng` provider in the reference's call order.

PINNING: tests/test_oracle_golden.py compares this against tests/golden/learner_*.pt recorded from
the unmodified reference (tests/golden/make_golden.py).  state_dict key names equal the reference's
so recorded parameters load directly.
"""
import torch
import torch.nn as nn

# AC_Args (actor_critic_decoder.py:11-88): the only hyper-dimensions the live path uses
TERRAIN_LATENT = 512
DIMS = dict(
    cenet_encoder=(265, [128], 64), cenet_decoder=(19 + TERRAIN_LATENT, [64, 128], 53),
    terrain_encoder=(693, [512, 512], TERRAIN_LATENT), terrain_decoder=(TERRAIN_LATENT, [512, 512], 693),
    memory_mlp=(265 + TERRAIN_LATENT, [256, 128], TERRAIN_LATENT), gb_encoder=(128, [128], 64),
)


def _ortho(layer, gain=0.01):
    nn.init.orthogonal_(layer.weight, gain)
    nn.init.constant_(layer.bias, 0.0)
    return layer


def mlp(inp, hidden, out, act):
    """First Linear keeps torch's default init, all later ones orthogonal(0.01)/zero bias; `act` between
    layers, none after the last (actor_critic_decoder.py:98-116 pattern)."""
    layers = [nn.Linear(inp, hidden[0]), act]
    for i in range(len(hidden)):
        if i == len(hidden) - 1:
            layers.append(_ortho(nn.Linear(hidden[i], out)))
        else:
            layers.append(_ortho(nn.Linear(hidden[i], hidden[i + 1])))
            layers.append(act)
    return nn.Sequential(*layers)


class Vae(nn.Module):
    def __init__(self, rng):
        super().__init__()
        self.rng = rng
        act = nn.ReLU()
        self.cenet_encoder = mlp(*DIMS["cenet_encoder"], act)
        self.latent_mu = _ortho(nn.Linear(64, 19))
        self.latent_var = _ortho(nn.Linear(64, 16))
        self.cenet_decoder = mlp(*DIMS["cenet_decoder"], act)
        self.terrain_encoder = mlp(*DIMS["terrain_encoder"], act)
        self.terrain_decoder = mlp(*DIMS["terrain_decoder"], act)
        self.memory_mlp = mlp(*DIMS["memory_mlp"], act)
        mlp(64, [128], 693, act)  # the reference builds and DISCARDS a ga_decoder here (:212-233); it consumes RNG
        self.gb_encoder = mlp(*DIMS["gb_encoder"], act)

    def cenet_forward(self, hist):
        e = self.cenet_encoder(hist)
        latent_var = self.latent_var(e)
        latent_mu = self.latent_mu(e)
        # outlier repair, actor_critic_decoder.py:293-299 (batch-global, in place)
        mean = latent_var.mean()
        std = latent_var.std()
        thr = 2 * std
        outliers = (latent_var < (mean - thr)) | (latent_var > (mean + thr))
        median = latent_var[~outliers].median()
        latent_var[outliers] = median
        stdv = torch.exp(0.5 * latent_var)
        eps = self.rng.randn_like(stdv)
        z = eps * stdv + latent_mu[:, 3:]
        return latent_mu, latent_var, z


class ActorCriticDecoder(nn.Module):
    is_recurrent = False

    def __init__(self, num_obs, num_critic_obs, num_actions, rng=None, **kwargs):
        super().__init__()
        self.rng = rng
        act = nn.ELU()
        self.vae = Vae(rng)
        self.actor_body = mlp(num_obs + 16 + 3 + TERRAIN_LATENT, [512, 256, 128], num_actions, act)
        self.critic_body = mlp(693 + num_obs + 3 + 15 + 12 - 24, [512, 256, 128], 1, act)
        self.std = nn.Parameter(1.0 * torch.ones(num_actions))
        self.distribution = None

    def reset(self, dones=None):
        pass

    @property
    def action_mean(self):
        return self.distribution.mean

    @property
    def action_std(self):
        return self.distribution.stddev

    @property
    def entropy(self):
        return self.distribution.entropy().sum(dim=-1)

    def update_distribution(self, obs, hist, priv):
        self.latent_mu, self.latent_var, self.z = self.vae.cenet_forward(hist)
        l_t = self.vae.terrain_encoder(priv[:, :693])
        mean = self.actor_body(torch.cat((obs, self.z, self.latent_mu[:, :3], l_t), dim=-1))
        self.distribution = torch.distributions.Normal(mean, mean * 0.0 + self.std, validate_args=False)

    def act(self, obs, hist, priv, rew_buf=None, **kw):
        self.update_distribution(obs, hist, priv)
        with torch.no_grad():
            d = self.distribution
            return d.mean + d.stddev * self.rng.randn_like(d.mean)

    def get_actions_log_prob(self, actions):
        return self.distribution.log_prob(actions).sum(dim=-1)

    def evaluate(self, obs, priv, base_vel, **kw):
        return self.critic_body(torch.cat((obs, base_vel, priv[:, 693:696], priv[:, 696:]), dim=-1))

    def act_teacher(self, obs, hist, priv):
        """Deployment path (actor_critic_decoder.py:504-538)."""
        latent = self.vae.latent_mu(self.vae.cenet_encoder(hist))
        l_t = self.vae.terrain_encoder(priv[:, :693])
        b_t1 = self.vae.memory_mlp(torch.cat((hist, l_t), dim=-1))
        b_t = b_t1 + torch.mul(l_t, b_t1)
        return self.actor_body(torch.cat((obs, latent[:, 3:], latent[:, :3], b_t), dim=-1))


class RolloutStorage:
    FIELDS = dict(observations=53, next_observations=53, privileged_observations=1389,
                  observation_histories=265, rewards=1, actions=12, actions_log_prob=1, values=1,
                  returns=1, advantages=1, mu=12, sigma=12, base_vel=3)

    class Transition:
        def __init__(self):
            self.observations = self.next_observations = self.privileged_observations = None
            self.observation_histories = self.critic_observations = self.actions = None
            self.rewards = self.dones = self.values = self.actions_log_prob = None
            self.action_mean = self.action_sigma = self.hidden_states = self.base_vel = None

        def clear(self):
            self.__init__()

    def __init__(self, num_envs, T, obs_shape, priv_shape, hist_shape, act_shape, device="cpu", rng=None):
        self.rng = rng
        self.num_envs, self.num_transitions_per_env = num_envs, T
        for k, d in self.FIELDS.items():
            setattr(self, k, torch.zeros(T, num_envs, d))
        self.dones = torch.zeros(T, num_envs, 1).byte()
        self.step = 0

    def add_transitions(self, t):
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        s = self.step
        self.observations[s].copy_(t.observations)
        self.next_observations[s].copy_(t.next_observations)
        self.privileged_observations[s].copy_(t.privileged_observations)
        self.observation_histories[s].copy_(t.observation_histories)
        self.actions[s].copy_(t.actions)
        self.rewards[s].copy_(t.rewards.view(-1, 1))
        self.dones[s].copy_(t.dones.view(-1, 1))
        self.values[s].copy_(t.values)
        self.actions_log_prob[s].copy_(t.actions_log_prob.view(-1, 1))
        self.mu[s].copy_(t.action_mean)
        self.base_vel[s].copy_(t.base_vel)
        self.sigma[s].copy_(t.action_sigma)
        self.step += 1

    def clear(self):
        self.step = 0

    def compute_returns(self, last_values, gamma, lam):
        advantage = 0
        for step in reversed(range(self.num_transitions_per_env)):
            nxt = last_values if step == self.num_transitions_per_env - 1 else self.values[step + 1]
            not_term = 1.0 - self.dones[step].float()
            delta = self.rewards[step] + not_term * gamma * nxt - self.values[step]
            advantage = delta + not_term * gamma * lam * advantage
            self.returns[step] = advantage + self.values[step]
        self.advantages = self.returns - self.values
        self.advantages = (self.advantages - self.advantages.mean()) / (self.advantages.std() + 1e-8)

    def mini_batch_generator(self, num_mini_batches, num_epochs=8):
        bs = self.num_envs * self.num_transitions_per_env
        mbs = bs // num_mini_batches
        indices = self.rng.randperm(num_mini_batches * mbs)
        f = {k: getattr(self, k).flatten(0, 1) for k in self.FIELDS}
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                b = indices[i * mbs:(i + 1) * mbs]
                yield (f["observations"][b], f["observations"][b], f["privileged_observations"][b],
                       f["observation_histories"][b], f["actions"][b], f["values"][b], f["advantages"][b],
                       f["returns"][b], f["actions_log_prob"][b], f["mu"][b], f["sigma"][b], f["base_vel"][b],
                       f["next_observations"][b], (None, None), None, f["rewards"][b])


class PPO:
    def __init__(self, actor_critic, num_learning_epochs=5, num_mini_batches=4, clip_param=0.2, gamma=0.99,
                 lam=0.95, value_loss_coef=1.0, entropy_coef=0.01, learning_rate=5.e-4, max_grad_norm=1.0,
                 use_clipped_value_loss=True, schedule="adaptive", desired_kl=0.01, device="cpu", rng=None):
        self.rng = rng
        self.desired_kl, self.schedule, self.learning_rate = desired_kl, schedule, learning_rate
        self.actor_critic = actor_critic
        self.storage = None
        self.optimizer = torch.optim.Adam(self.actor_critic.parameters(), lr=learning_rate)
        self.vae_optimizer = torch.optim.Adam(self.actor_critic.vae.parameters(), lr=5.e-4)
        self.transition = RolloutStorage.Transition()
        self.clip_param, self.num_learning_epochs, self.num_mini_batches = clip_param, num_learning_epochs, num_mini_batches
        self.value_loss_coef, self.entropy_coef = value_loss_coef, entropy_coef
        self.gamma, self.lam, self.max_grad_norm = gamma, lam, max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self.debug = None  # optional dict collecting per-minibatch internals for kernel tests

    def init_storage(self, num_envs, T, obs_shape, priv_shape, hist_shape, act_shape):
        self.storage = RolloutStorage(num_envs, T, obs_shape, priv_shape, hist_shape, act_shape, rng=self.rng)

    def act(self, obs, priv, hist, base_vel, rew_buf=None):
        t = self.transition
        t.actions = self.actor_critic.act(obs, hist, priv, rew_buf).detach()
        t.values = self.actor_critic.evaluate(obs, priv, base_vel).detach()
        t.actions_log_prob = self.actor_critic.get_actions_log_prob(t.actions).detach()
        t.action_mean = self.actor_critic.action_mean.detach()
        t.action_sigma = self.actor_critic.action_std.detach()
        t.observations = t.critic_observations = obs
        t.privileged_observations, t.observation_histories, t.base_vel = priv, hist, base_vel
        return t.actions

    def process_env_step(self, rewards, dones, next_obs, infos):
        t = self.transition
        t.rewards = rewards.clone()
        t.dones = dones
        t.next_observations = next_obs
        if "time_outs" in infos:
            t.rewards += self.gamma * torch.squeeze(t.values * infos["time_outs"].unsqueeze(1), 1)
        self.storage.add_transitions(t)
        t.clear()

    def compute_returns(self, obs, priv, base_vel):
        last_values = self.actor_critic.evaluate(obs, priv, base_vel).detach()
        self.storage.compute_returns(last_values, self.gamma, self.lam)

    def update(self):
        ac = self.actor_critic
        m_value = m_surr = m_recons = m_vel = m_kld = 0.0
        for (obs_b, critic_obs_b, priv_b, hist_b, actions_b, target_values_b, adv_b, returns_b, old_logp_b,
             old_mu_b, old_sigma_b, base_vel_b, next_obs_b, _, _, rew_b) in self.storage.mini_batch_generator(
                self.num_mini_batches, self.num_learning_epochs):
            # ---- VAE step (ppo.py:197-254)
            latent_mu, latent_var, z = ac.vae.cenet_forward(hist_b)
            l_t = ac.vae.terrain_encoder(priv_b[:, :693])
            recons = ac.vae.cenet_decoder(torch.cat([z, latent_mu[:, :3], l_t], dim=1))
            recons_loss = torch.pow(recons - next_obs_b, 2).mean(-1).mean()
            height_recon = ac.vae.terrain_decoder(l_t)
            height_loss = torch.nn.functional.mse_loss(height_recon, priv_b[..., 696:])
            vel_loss = torch.nn.functional.mse_loss(latent_mu[:, :3], base_vel_b)
            kld_loss = torch.mean(-0.5 * torch.sum(1 + latent_var - latent_mu[:, 3:].pow(2) - latent_var.exp(), dim=1))
            vae_loss = recons_loss + vel_loss + 4 * kld_loss + height_loss
            self.vae_optimizer.zero_grad()
            vae_loss.backward()
            if self.debug is not None:
                self.debug.setdefault("vae_grads", []).append(
                    {k: p.grad.clone() for k, p in ac.vae.named_parameters() if p.grad is not None})
                self.debug.setdefault("vae_losses", []).append(
                    (recons_loss.item(), vel_loss.item(), kld_loss.item(), height_loss.item()))
            nn.utils.clip_grad_norm_(ac.vae.parameters(), self.max_grad_norm)
            self.vae_optimizer.step()
            m_recons += recons_loss.item()
            m_vel += vel_loss.item()
            m_kld += kld_loss.item()
            # ---- policy step (ppo.py:265-338)
            ac.act(obs_b, hist_b, priv_b, rew_b)
            logp_b = ac.get_actions_log_prob(actions_b)
            value_b = ac.evaluate(critic_obs_b, priv_b, base_vel_b)
            mu_b, sigma_b, entropy_b = ac.action_mean, ac.action_std, ac.entropy
            if self.desired_kl is not None and self.schedule == "adaptive":
                with torch.inference_mode():
                    kl = torch.sum(torch.log(sigma_b / old_sigma_b + 1.e-5) + (torch.square(old_sigma_b) + torch.square(
                        old_mu_b - mu_b)) / (2.0 * torch.square(sigma_b)) - 0.5, axis=-1)
                    kl_mean = torch.mean(kl)
                    if kl_mean > self.desired_kl * 2.0:
                        self.learning_rate = max(1e-5, self.learning_rate / 1.5)
                    elif kl_mean < self.desired_kl / 2.0 and kl_mean > 0.0:
                        self.learning_rate = min(1e-2, self.learning_rate * 1.5)
                    for g in self.optimizer.param_groups:
                        g["lr"] = self.learning_rate
            ratio = torch.exp(logp_b - torch.squeeze(old_logp_b))
            surrogate = -torch.squeeze(adv_b) * ratio
            surrogate_clipped = -torch.squeeze(adv_b) * torch.clamp(ratio, 1.0 - self.clip_param, 1.0 + self.clip_param)
            surrogate_loss = torch.max(surrogate, surrogate_clipped).mean()
            if self.use_clipped_value_loss:
                value_clipped = target_values_b + (value_b - target_values_b).clamp(-self.clip_param, self.clip_param)
                value_loss = torch.max((value_b - returns_b).pow(2), (value_clipped - returns_b).pow(2)).mean()
            else:
                value_loss = (returns_b - value_b).pow(2).mean()
            loss = surrogate_loss + self.value_loss_coef * value_loss - self.entropy_coef * entropy_b.mean()
            self.optimizer.zero_grad()
            loss.backward()
            if self.debug is not None:
                self.debug.setdefault("ppo_grads", []).append(
                    {k: p.grad.clone() for k, p in ac.named_parameters() if p.grad is not None})
                self.debug.setdefault("ppo_losses", []).append(
                    (value_loss.item(), surrogate_loss.item(), entropy_b.mean().item(), float(kl_mean), self.learning_rate))
            nn.utils.clip_grad_norm_(ac.parameters(), self.max_grad_norm)
            self.optimizer.step()
            m_value += value_loss.item()
            m_surr += surrogate_loss.item()
        n = self.num_learning_epochs * self.num_mini_batches
        self.storage.clear()
        return m_value / n, m_surr / n, 0.0, 0, m_recons / n, m_vel / n, m_kld / n


def learn_iteration(env_wrapped, alg, obs_dict, T):
    """One iteration of OnPolicyRunner.learn (on_policy_runner.py:112-151) without logging/checkpoints.
    Returns (new obs_dict, update() tuple)."""
    obs, priv, hist = obs_dict["obs"], obs_dict["privileged_obs"], obs_dict["obs_history"]
    rew_buf = env_wrapped.get_reward_buf()
    with torch.inference_mode():
        for _ in range(T):
            actions = alg.act(obs, priv, hist, obs_dict["base_vel"], rew_buf)
            obs_dict, rewards, dones, infos = env_wrapped.step(actions)
            obs, priv, hist = obs_dict["obs"], obs_dict["privileged_obs"], obs_dict["obs_history"]
            alg.process_env_step(rewards, dones, next_obs=obs_dict["obs"], infos=infos)
        alg.compute_returns(obs, priv, obs_dict["base_vel"])
    return obs_dict, alg.update()
