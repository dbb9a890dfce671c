"""Genetic algorithm over the five runtime MPC parameters (GA/mpcGA.py), with the fitness the
reference leaves as a placeholder (``segment_time = np.random.rand()``, GA/mpcGA.py:43) supplied
by batched closed-loop MPC rollouts on the GPU.

Constants and reward/EMA formulas are the reference's (GA/mpcGA.py:6-24); selection, one-point
crossover and mutation follow the working template GA/pdGA.py:57-91.  The reference never
defines how a chromosome in [0,1)^5 maps to RuntimeControllerParameters: PARAM_RANGES below is
this build's declaration of that map.
"""
import numpy as np

population_size = 100    # Npop
num_parameters = 5       # (alpha_c, d_max, q_v_y, n, beta_delta)
num_generations = 10
num_checkpoints = 5      # NCP
mutation_rate = 0.4      # P(M)
crossover_rate = 0.5     # P(C)
kT = 4                   # s^-1
kc = 0.5
lambda_T = 0.3

# chromosome gene in [0,1) -> parameter value (linear; n snaps to the even exponents 2 / 4)
PARAM_RANGES = dict(alpha_c=(100.0, 2000.0), d_max=(0.5, 1.0), q_v_y=(0.0, 100.0), n=(2, 4),
                    beta_delta=(500.0, 10000.0))


def decode(chromosomes):
    """[P,5] genes in [0,1) -> [P,5] rows (alpha_c, d_max, q_v_y, n, beta_delta)."""
    g = np.clip(np.asarray(chromosomes, dtype=np.float64), 0.0, 1.0 - 1e-12)
    out = np.empty_like(g)
    for j, key in enumerate(("alpha_c", "d_max", "q_v_y", "n", "beta_delta")):
        lo, hi = PARAM_RANGES[key]
        out[:, j] = lo + g[:, j] * (hi - lo)
    out[:, 3] = np.where(g[:, 3] < 0.5, 2.0, 4.0)
    return out


def initialize_population(population_size, num_parameters, rng=np.random):
    return rng.rand(population_size, num_parameters)


def compute_reward(segment_time, average_time):
    return np.exp(-kT * (segment_time - average_time))


def compute_avg_time(segment_time, average_time):
    return lambda_T * segment_time + (1 - lambda_T) * average_time


class GeneticAlgorithm:
    """population x checkpoints fitness from ``rollouts`` (a ClosedLoopRollouts)."""

    def __init__(self, rollouts, seg_bounds, lap_time_guess, pop=population_size, seed=1234):
        self.ro = rollouts
        self.seg_bounds = list(seg_bounds)
        self.n_cp = len(self.seg_bounds) - 1
        self.avg_times = np.full(self.n_cp, lap_time_guess / self.n_cp)
        self.rng = np.random.RandomState(seed)
        self.population = initialize_population(pop, num_parameters, self.rng)

    def evaluate_population(self, x0, s0, max_steps):
        P = len(self.population)
        prm = decode(self.population)
        res = self.ro.run(np.repeat(x0[None], P, 0), np.repeat(s0, P), params=prm, max_steps=max_steps,
                          seg_bounds=self.seg_bounds)
        seg = res["seg_times"]                                     # [P, n_cp], nan = not reached
        rewards = np.zeros(P)
        for i in range(P):
            for c in range(self.n_cp):
                if np.isnan(seg[i, c]):
                    continue                                       # unfinished segment: no reward
                average_time = self.avg_times[c]
                self.avg_times[c] = compute_avg_time(seg[i, c], average_time)
                rewards[i] += compute_reward(seg[i, c], average_time)
        return rewards, res

    def select_parents(self, rewards):
        probs = rewards / np.sum(rewards) if np.sum(rewards) > 0 else np.full(len(rewards), 1.0 / len(rewards))
        idx = self.rng.choice(np.arange(len(rewards)), size=2, p=probs)
        return self.population[idx]

    def crossover(self, p1, p2):
        cp = self.rng.randint(1, num_parameters)
        return np.concatenate((p1[:cp], p2[cp:])), np.concatenate((p2[:cp], p1[cp:]))

    def mutate(self, child):
        m = self.rng.rand(num_parameters) < mutation_rate
        child = child.copy()
        child[m] += (self.rng.rand(num_parameters) - 0.5)[m]
        return np.clip(child, 0.0, 1.0 - 1e-12)

    def run_step(self, x0, s0, max_steps=400):
        rewards, res = self.evaluate_population(x0, s0, max_steps)
        p1, p2 = self.select_parents(rewards)
        if self.rng.rand() < crossover_rate:
            c1, c2 = self.crossover(p1, p2)
        else:
            c1, c2 = p1.copy(), p2.copy()
        worst = np.argsort(rewards)[:2]
        self.population[worst[0]] = self.mutate(c1)
        self.population[worst[1]] = self.mutate(c2)
        best = int(np.argmax(rewards))
        return dict(best_index=best, best_params=decode(self.population[best:best + 1])[0],
                    best_reward=float(rewards[best]), rewards=rewards, rollouts=res)


class MultiTrackGA(GeneticAlgorithm):
    """BASELINE config 5: the population's fitness over SEVERAL tracks, with the (track, individual)
    rollouts sharded over the ranks of one box (one process per GPU) and only the small fitness
    table - segment times [tracks x population, checkpoints] - exchanged (one all-gather per
    generation, NCCL when ``dist`` is an initialised torch.distributed with CUDA tensors).

    ``courses``: list of dicts, one per track: ``ro`` (ClosedLoopRollouts of that track, or None on
    ranks that own no rollout of it), ``seg_bounds``, ``lap_time_guess``, ``x0`` [6], ``s0``.
    Rewards follow GA/mpcGA.py:35-48 track by track: every (track, checkpoint) keeps its own
    running average time, an individual's reward is the sum over tracks and checkpoints.
    """

    def __init__(self, courses, pop=population_size, seed=1234, rank=0, world=1, dist=None, device=None):
        self.courses = courses
        self.rank, self.world, self.dist, self.device = rank, world, dist, device
        self.n_cp = len(courses[0]["seg_bounds"]) - 1
        self.avg_times = np.stack([np.full(self.n_cp, c["lap_time_guess"] / self.n_cp) for c in courses])
        self.rng = np.random.RandomState(seed)
        self.population = initialize_population(pop, num_parameters, self.rng)

    def evaluate_population(self, max_steps):
        from ..sharding import shard_range, gather_results
        P, T = len(self.population), len(self.courses)
        prm = decode(self.population)
        lo, hi = shard_range(T * P, self.rank, self.world)           # pair index = track * P + individual
        seg_local = np.full((hi - lo, self.n_cp), np.nan)
        fails_local = np.zeros(hi - lo, dtype=np.int32)
        for t, c in enumerate(self.courses):
            a, b = max(lo, t * P), min(hi, (t + 1) * P)
            if a >= b:
                continue
            ind = np.arange(a - t * P, b - t * P)
            res = c["ro"].run(np.repeat(np.asarray(c["x0"])[None], len(ind), 0), np.repeat(c["s0"], len(ind)),
                              params=prm[ind], max_steps=max_steps, seg_bounds=c["seg_bounds"])
            seg_local[a - lo:b - lo] = res["seg_times"]
            fails_local[a - lo:b - lo] = res["fail_steps"]
        full = gather_results(dict(seg=seg_local, fails=fails_local), T * P, self.rank, self.world, self.dist,
                              self.device)
        seg = full["seg"].reshape(T, P, self.n_cp)
        rewards = np.zeros(P)
        for t in range(T):
            for i in range(P):
                for cpt in range(self.n_cp):
                    if np.isnan(seg[t, i, cpt]):
                        continue                                   # unfinished segment: no reward
                    average_time = self.avg_times[t, cpt]
                    self.avg_times[t, cpt] = compute_avg_time(seg[t, i, cpt], average_time)
                    rewards[i] += compute_reward(seg[t, i, cpt], average_time)
        return rewards, dict(seg_times=seg, fail_steps=full["fails"].reshape(T, P))

    def run_step(self, max_steps=400):
        rewards, res = self.evaluate_population(max_steps)
        # every rank holds the same rewards and the same random stream: identical populations afterwards
        p1, p2 = self.select_parents(rewards)
        if self.rng.rand() < crossover_rate:
            c1, c2 = self.crossover(p1, p2)
        else:
            c1, c2 = p1.copy(), p2.copy()
        worst = np.argsort(rewards)[:2]
        self.population[worst[0]] = self.mutate(c1)
        self.population[worst[1]] = self.mutate(c2)
        best = int(np.argmax(rewards))
        return dict(best_index=best, best_params=decode(self.population[best:best + 1])[0],
                    best_reward=float(rewards[best]), rewards=rewards, rollouts=res)
