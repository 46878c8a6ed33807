"""Mirror of base/core/agent.py (Agent :13-352): same constructor, evaluate(), train(), validate_agent(), save_agent() and
stats keys, with the per-generation fitness hot path on the GPU:

  * the population loop `for net in pop: for i in range(num_evals): evaluate(net)` (:234-241) is ONE fused rollout launch
    over pop x num_envs trajectories (serl_b200/rollout.py, csrc/rollout.cu), sharded over ranks when torch.distributed
    is initialised (serl_b200/engine.py);
  * single episodes (`evaluate`, used for the champion / RL validation and the RL exploration episode) run the same
    kernel with a per-step trace, from which the Episode record and the replay transitions are rebuilt;
  * `self.evolver.epoch` runs on the device-resident genomes (core/mod_neuro_evo.py -> serl_b200/evo.py).

Documented deviations from the reference: the conditions at agent.py:45,228 are read as intended
(`if self.pop`), save_agent's `isEmpty()` works; all actors of a generation see the SAME num_envs reference signals (fair
ranking) instead of an independent draw per episode; every trajectory starts from a fresh env (zero stale error).
"""
import os
from typing import Dict

import numpy as np
import torch

from . import mod_utils, replay_memory, td3
from . import mod_neuro_evo as utils_ne
from .utils import Episode, calc_smoothness
from .. import engine, refsig, rollout
from ..population import PopulationList


class Agent:
    def __init__(self, args, environment):
        self.args = args
        self.env = environment
        if not torch.cuda.is_available():
            from .._native import NativeError
            raise NativeError('serl_b200.Agent needs a CUDA device: the rollout / evolution engine has no CPU fallback')
        self.device = torch.device('cuda', torch.cuda.current_device())
        self.pop = PopulationList(args, self.device) if args.pop_size else []
        self.rl_agent = td3.TD3(args)
        self.replay_buffer = replay_memory.ReplayMemory(args.buffer_size, args.device)
        self.noise_process = mod_utils.GaussianNoise(args.action_dim, sd=args.noise_sd)
        if len(self.pop):
            self.evolver = utils_ne.SSNE(self.args, self.rl_agent.critic, self.evaluate)
        self.shape = rollout.actor_shape(args.hidden_size, args.num_layers, args.activation_actor, args.state_dim, args.action_dim)
        self.validation_tests = 5
        self.num_episodes = 0
        self.num_frames = 0
        self.iterations = 0
        self.gen_frames = None
        self.rl_history = None
        self.rl_iteration = 0
        self.champion = None
        self.champion_actor = None
        self.champion_history = None
        self.store_population_transitions = args.frac_frames_train > 0

    # ------------------------------------------------------------------------------------------------ episodes
    def _genome_of(self, agent):
        idx = getattr(agent, 'index', None)
        if idx is not None and len(self.pop) and self.pop[idx] is agent:
            return self.pop.genomes[idx:idx + 1]
        return agent.actor.flat().to(self.device).reshape(1, -1).contiguous()

    def _final_time(self, n):
        t = 0.
        for _ in range(n):
            t += self.env.dt
        return t

    def _episode_from_trace(self, agent, tr, n, x_ic, store_transition, refs):
        rewards = [float(r) for r in tr[:n, 15]]
        actions = tr[:n, 12:15].copy()
        if store_transition:
            # transitions (obs, action, next_obs, reward, done) of agent.py:101-112, rebuilt from the trace in one shot
            x = tr[:n, 0:12]
            next_obs = np.hstack((tr[:n, 19:22], x[:, [0, 1, 2, 4]]))
            obs = np.vstack((np.hstack((np.zeros(3), x_ic[[0, 1, 2, 4]]))[None], next_obs[:-1]))
            done = np.zeros(n); done[-1] = 1.0
            rew = np.asarray(rewards)
            batch = (obs, tr[:n, 16:19], next_obs, rew, done)
            self.replay_buffer.add_batch(*batch)
            agent.buffer.add_batch(*batch)
            # get_cost (phlabenv.py:369-375, incl. its degrees-vs-radians comparison on the bank angle)
            cost = (np.rad2deg(np.abs(x[:, 4])) > 11.0) | (np.rad2deg(np.abs(x[:, 6])) > 0.75 * self.env.max_phi) | (x[:, 3] < x_ic[3] / 3)
            if cost.any():
                agent.critical_buffer.add_batch(*(b[cost] for b in batch))
            self.num_frames += n
            self.gen_frames += n
            self.num_episodes += 1
            state_lst = []
        else:
            state_lst = [tr[k, 0:12].copy() for k in range(n)]
        smoothness = calc_smoothness(actions, plot_spectra=False)
        fitness = np.sum(rewards)
        if self.args.smooth_fitness:
            fitness += smoothness
        return Episode(fitness=fitness, smoothness=smoothness, length=self._final_time(n), state_history=state_lst,
                       ref_signals=refs, actions=actions, reward_lst=rewards)

    def evaluate(self, agent, is_action_noise: bool, store_transition: bool) -> Episode:
        """Play one episode (agent.py:63-138) on the GPU."""
        env = self.env
        levels, starts = env.draw_reference()
        lv = torch.as_tensor(levels[None], device=self.device)
        st = torch.as_tensor(starts[None], device=self.device)
        md = torch.tensor([env.mode_code], dtype=torch.int32, device=self.device)
        noise = None
        if is_action_noise:
            # one np.random.randn(3) per executed step (agent.py:90-93): draw a full horizon, then rewind the global
            # stream to "exactly the steps that ran" once the episode length is known
            state = np.random.get_state()
            z = np.random.randn(int(round(env.t_max / env.dt)) + 1, 3)
            clipped = np.clip(self.args.noise_sd * z, -self.args.noise_clip, self.args.noise_clip)
            noise = torch.as_tensor(clipped.astype(np.float32).reshape(1, 1, -1, 3), device=self.device)
        horizon = int(round(env.t_max / env.dt)) + 1
        kw = {} if env.t_max == 20 else {'t_max': float(env.t_max), 'smooth_width': refsig.widths(env.t_max)[1]}
        r = rollout.population_rollout(self._genome_of(agent), self.shape, lv, st, md, trace=True, action_noise=noise,
                                       horizon=horizon, **kw)
        n = int(r.steps[0, 0].item())
        tr = r.trace[0, 0, :n].cpu().numpy()
        if is_action_noise:
            np.random.set_state(state)
            np.random.randn(n, 3)
        x_ic = self._initial_state(env)
        theta_trim = np.rad2deg(x_ic[7])
        from ..envs.phlabenv import _RefSignal
        sw = refsig.widths(env.t_max)[1]
        refs = [_RefSignal(levels[0], starts[0], theta_trim, sw, env.t_max), _RefSignal(levels[1], starts[1], 0.0, sw), lambda t: 0.0]
        return self._episode_from_trace(agent, tr, n, x_ic, store_transition, refs)

    _ic_cache: Dict[int, np.ndarray] = {}

    def _initial_state(self, env):
        v = env.mode_code & 0xff
        if v not in Agent._ic_cache:
            import ctypes
            from .. import _native
            X = torch.empty((1, 19), dtype=torch.float64, device=self.device)
            var = torch.tensor([v], dtype=torch.int32, device=self.device)
            _native.check(_native.lib().serl_plant_init(ctypes.c_void_p(X.data_ptr()), ctypes.c_void_p(var.data_ptr()), 1,
                                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), 'serl_plant_init')
            Agent._ic_cache[v] = X.cpu().numpy()[0, :12].copy()
        return Agent._ic_cache[v]

    def rl_to_evo(self, rl_agent, evo_net):
        for target_param, param in zip(evo_net.actor.parameters(), rl_agent.actor.parameters()):
            target_param.data.copy_(param.data)
        evo_net.buffer.reset()
        evo_net.buffer.add_content_of(rl_agent.buffer)
        evo_net.critical_buffer.reset()
        evo_net.critical_buffer.add_content_of(rl_agent.critical_buffer)

    def evo_to_rl(self, rl_net, evo_net):
        for target_param, param in zip(rl_net.parameters(), evo_net.parameters()):
            target_param.data.copy_(param.data)

    def train_rl(self, rl_transitions: int) -> Dict[str, float]:
        pgs_obj, TD_loss = [], []
        if len(self.replay_buffer) > self.args.learn_start:
            self.rl_agent.actor.train()
            if self.args.use_champion_target and self.champion_actor is not None:
                self.evo_to_rl(self.rl_agent.actor_target, self.champion_actor)
            for _ in range(int(rl_transitions * self.args.frac_frames_train)):
                self.rl_iteration += 1
                batch = self.replay_buffer.sample(self.args.batch_size)
                pgl, TD = self.rl_agent.update_parameters(batch, self.rl_iteration, self.args.use_champion_target)
                if pgl is not None:
                    pgs_obj.append(-pgl)
                if TD is not None:
                    TD_loss.append(TD)
        return {'PG_obj': np.mean(pgs_obj) if pgs_obj else float('nan'), 'TD_loss': np.median(TD_loss) if TD_loss else float('nan')}

    def validate_agent(self, agent):
        test_scores, episode_lengths, smoothness_lst = [], [], []
        for _ in range(self.validation_tests):
            last_episode = self.evaluate(agent, is_action_noise=False, store_transition=False)
            test_scores.append(np.sum(last_episode.reward_lst))
            episode_lengths.append(last_episode.length)
            smoothness_lst.append(last_episode.smoothness)
        return (np.mean(test_scores), np.std(test_scores), np.mean(episode_lengths), np.std(episode_lengths), last_episode,
                np.median(smoothness_lst), np.std(smoothness_lst))

    # ------------------------------------------------------------------------------------------------ generation
    def evaluate_population(self):
        """agent.py:229-245 as one fused launch. Returns (pop_fitness f64[pop] numpy, lengths list, device fitness)."""
        n_envs = int(getattr(self.args, 'num_envs', self.args.num_evals))
        self._count_before, self._gen_before = (self.num_frames, self.num_episodes), self.gen_frames
        draws = [self.env.draw_reference() for _ in range(n_envs)]
        lv = torch.as_tensor(np.stack([d[0] for d in draws]), device=self.device)
        st = torch.as_tensor(np.stack([d[1] for d in draws]), device=self.device)
        md = torch.full((n_envs,), self.env.mode_code, dtype=torch.int32, device=self.device)
        want_sm = bool(getattr(self.args, 'population_smoothness', True))
        fitness, r, (lo, hi) = engine.evaluate_population(self.pop.genomes, self.shape, lv, st, md, actions=want_sm,
                                                          smooth_fitness=bool(self.args.smooth_fitness))
        steps = r.steps.cpu().numpy() if r is not None else np.zeros((0, n_envs), dtype=np.int32)
        self._last_smoothness = r.smoothness.cpu().numpy() if (r is not None and getattr(r, 'smoothness', None) is not None) else None
        lengths = [self._final_time(int(s)) for s in steps.reshape(-1)[:256]]     # statistic only; bounded host work
        if self.store_population_transitions and hi > lo:
            # transitions of the last evaluation of every actor (agent.py:236-238), from a traced re-flight of that env
            tr = rollout.population_rollout(self.pop.genomes[lo:hi], self.shape, lv[-1:].contiguous(), st[-1:].contiguous(), md[-1:], trace=True)
            x_ic = self._initial_state(self.env)
            tsteps = tr.steps[:, 0].cpu().numpy()
            for a in range(hi - lo):
                n = int(tsteps[a])
                self._episode_from_trace(self.pop[lo + a], tr.trace[a, 0, :n].cpu().numpy(), n, x_ic, True, None)
        else:
            self.num_frames += int(steps[:, -1].sum()) if steps.size else 0
            self.gen_frames += int(steps[:, -1].sum()) if steps.size else 0
            self.num_episodes += hi - lo
        world, _ = engine.world_info()
        if world > 1:
            # frame / episode counters drive the training loop (base/train.py:102); keep them identical on every rank
            import torch.distributed as dist
            f0, e0 = self._count_before
            cnt = torch.tensor([self.num_frames - f0, self.num_episodes - e0, self.gen_frames - self._gen_before],
                               dtype=torch.int64, device=self.device)
            dist.all_reduce(cnt)
            self.num_frames, self.num_episodes = f0 + int(cnt[0]), e0 + int(cnt[1])
            self.gen_frames = self._gen_before + int(cnt[2])
        return fitness.cpu().numpy(), lengths, fitness

    def train(self):
        self.iterations += 1
        self.gen_frames = 0
        best_train_fitness = worst_train_fitness = population_avg = test_score = sm = 1.
        test_sd = sm_sd = elite_index = pop_novelty = -1.
        ep_len_avg = ep_len_sd = 0.
        pop_fitness = None
        if len(self.pop):
            pop_fitness, lengths, dev_fitness = self.evaluate_population()
            if self._last_smoothness is not None:      # K6: per-episode action smoothness on the device (agent.py:242-243)
                sm, sm_sd = float(np.mean(self._last_smoothness)), float(np.std(self._last_smoothness))
            else:
                sm, sm_sd = float('nan'), float('nan')
            ep_len_avg, ep_len_sd = np.mean(lengths), np.std(lengths)
            best_train_fitness = np.max(pop_fitness)
            worst_train_fitness = np.min(pop_fitness)
            population_avg = np.average(pop_fitness)
            self.champion = self.pop[int(np.argmax(pop_fitness))]
            self.champion_actor = self.champion.actor
            test_score, test_sd, _, _, last_episode, _, _ = self.validate_agent(self.champion)
            if self.args.should_log:
                self.champion_history = last_episode.get_history()
            elite_index = self.evolver.epoch(self.pop, dev_fitness)
        # RL half (agent.py:267-281)
        self.evaluate(self.rl_agent, is_action_noise=True, store_transition=True)
        rl_train_scores = self.train_rl(self.gen_frames)
        rl_reward, rl_std, rl_ep_len, rl_ep_std, rl_episode, rl_sm, rl_sm_sd = self.validate_agent(self.rl_agent)
        if self.args.pop_size == 0:
            ep_len_avg, ep_len_sd = rl_ep_len, rl_ep_std
        if self.args.should_log:
            self.rl_history = rl_episode.get_history()
        # actor injection (agent.py:283-294)
        if self.args.pop_size and self.iterations % self.args.rl_to_ea_synch_period == 0:
            replace_index = int(np.argmin(pop_fitness))
            if replace_index == elite_index:
                replace_index = (replace_index + 1) % len(self.pop)
            self.rl_to_evo(self.rl_agent, self.pop[replace_index])
            self.evolver.rl_policy = replace_index
        return {
            'best_train_fitness': best_train_fitness, 'test_score': test_score, 'test_sd': test_sd,
            'pop_avg': population_avg, 'pop_min': worst_train_fitness, 'elite_index': elite_index,
            'avg_smoothness': sm, 'smoothness_sd': sm_sd, 'rl_reward': rl_reward, 'rl_smoothness': rl_sm,
            'rl_smoothness_std': rl_sm_sd, 'rl_std': rl_std, 'avg_ep_len': ep_len_avg, 'ep_len_sd': ep_len_sd,
            'PG_obj': rl_train_scores['PG_obj'], 'TD_loss': rl_train_scores['TD_loss'], 'pop_novelty': pop_novelty,
        }

    def save_agent(self, parameters, elite_index: int = None) -> None:
        """agent.py:317-352: evo_nets.pkl ({'actor_i': state_dict}), elite_net.pkl, rl_net.pkl, state histories."""
        if len(self.pop):
            pop_dict = {f'actor_{i}': {k: v.detach().cpu().clone() for k, v in ind.actor.state_dict().items()}
                        for i, ind in enumerate(self.pop)}
            torch.save(pop_dict, os.path.join(parameters.save_foldername, 'evo_nets.pkl'))
            torch.save(pop_dict[f'actor_{int(elite_index)}'], os.path.join(parameters.save_foldername, 'elite_net.pkl'))
            if self.champion_history is not None:
                np.savetxt(os.path.join(parameters.save_foldername, 'statehistory_episode%d.txt' % self.num_episodes),
                           self.champion_history, header=str(self.num_episodes))
        torch.save({k: v.detach().cpu() for k, v in self.rl_agent.actor.state_dict().items()},
                   os.path.join(parameters.save_foldername, 'rl_net.pkl'))
        if self.rl_history is not None:
            np.savetxt(os.path.join(parameters.save_foldername, 'rl_statehistory_episode%d.txt' % self.num_episodes),
                       self.rl_history, header=str(self.num_episodes))
