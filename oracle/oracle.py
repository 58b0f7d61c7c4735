"""ctypes front-end of the C oracle (oracle/orlg_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liborlg_oracle.so")

KINDS = {"RWA-v0": 0, "RMSA-v0": 1, "DeepRMSA-v0": 2, "RMCSA-v0": 3}
ACTION_DIM = {0: 2, 1: 2, 2: 1, 3: 4}


def build(force=False):
    src = os.path.join(HERE, "orlg_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "liborlg_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB


class Cfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "kind", "num_nodes", "num_links", "k_paths", "num_paths", "num_slots", "num_cores", "num_mods", "j",
        "episode_length", "allow_rejection", "bit_rate_lo", "bit_rate_hi", "num_bit_rates", "stats")] + [
        (n, C.c_double) for n in ("channel_width", "mean_holding", "mean_iat", "worst_xt")]


class Tab(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "pair_first", "pair_count", "path_hops", "path_se", "path_mod", "path_link_ptr", "path_links",
        "path_length", "mod_se", "mod_osnr", "mod_xt", "node_prob", "bit_rates", "bit_rate_prob", "link_order")]


class StepOut(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "accepted", "path_row", "initial_slot", "number_slots", "core", "mod", "service_id", "done")] + [
        ("reward", C.c_double), ("info_counters", C.c_int64 * 8), ("stats", C.c_double * 4)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(Cfg), C.POINTER(Tab)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_trace.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int64]
        L.oracle_set_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        L.oracle_reset.argtypes = [C.c_void_p, C.c_int]
        L.oracle_step.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(StepOut)]
        L.oracle_step.restype = C.c_int
        L.oracle_observation.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_observation_int.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_heuristic.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_random_action.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_get_request.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_get_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_get_bit_rate_blocking.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_get_action_probability.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_get_link_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_get_link_stats.restype = None
        L.oracle_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.oracle_error.argtypes = [C.c_void_p]
        L.oracle_error.restype = C.c_int
        L.oracle_rollout.restype = C.c_long
        L.oracle_rollout.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_rollout_mt.restype = C.c_long
        L.oracle_rollout_mt.argtypes = [C.POINTER(Cfg), C.POINTER(Tab), C.c_uint64, C.c_uint32, C.c_int, C.c_long,
                                        C.c_int, C.c_int, C.c_int]
        L.oracle_vec_create.restype = C.c_void_p
        L.oracle_vec_create.argtypes = [C.POINTER(Cfg), C.POINTER(Tab), C.c_uint64, C.c_uint32, C.c_int]
        L.oracle_vec_destroy.argtypes = [C.c_void_p]
        L.oracle_vec_run.restype = C.c_long
        L.oracle_vec_run.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int]
        L.oracle_vec_env.restype = C.c_void_p
        L.oracle_vec_env.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_cfg_tab(env_id, tables, *, num_slots, episode_length=1000, j=1, num_cores=1, allow_rejection=False,
                 mean_holding=25.0, mean_iat=0.1, channel_width=12.5, bit_rate_lo=25, bit_rate_hi=100,
                 bit_rates=None, bit_rate_prob=None, node_prob=None, worst_xt=-84.7, stats=True):
    """Returns (Cfg, Tab, keepalive list of numpy arrays)."""
    t = tables
    keep = {}

    def arr(name, a, dt):
        keep[name] = np.ascontiguousarray(a, dtype=dt)
        return _ptr(keep[name])

    if node_prob is None:
        node_prob = np.full(t.num_nodes, 1.0 / t.num_nodes)
    nbr = 0 if bit_rates is None else len(bit_rates)
    if nbr and bit_rate_prob is None:
        bit_rate_prob = [1.0 / nbr] * nbr
    cfg = Cfg(kind=KINDS[env_id], num_nodes=t.num_nodes, num_links=t.num_links, k_paths=t.k_paths,
              num_paths=t.num_paths, num_slots=num_slots, num_cores=num_cores, num_mods=len(t.mod_se), j=j,
              episode_length=episode_length, allow_rejection=int(allow_rejection), bit_rate_lo=int(bit_rate_lo),
              bit_rate_hi=int(bit_rate_hi), num_bit_rates=nbr, stats=int(stats), channel_width=channel_width,
              mean_holding=mean_holding, mean_iat=mean_iat, worst_xt=worst_xt)
    tab = Tab(pair_first=arr("pf", t.pair_first, np.int32), pair_count=arr("pc", t.pair_count, np.int32),
              path_hops=arr("ph", t.path_hops, np.int32), path_se=arr("ps", t.path_se, np.int32),
              path_mod=arr("pm", t.path_mod, np.int32), path_link_ptr=arr("pp", t.path_link_ptr, np.int32),
              path_links=arr("pl", t.path_links, np.int32), path_length=arr("plen", t.path_length, np.float64),
              mod_se=arr("ms", t.mod_se, np.int32), mod_osnr=arr("mo", t.mod_osnr, np.float64),
              mod_xt=arr("mx", t.mod_xt, np.float64), node_prob=arr("np", node_prob, np.float64),
              bit_rates=arr("br", bit_rates if nbr else [0], np.int32),
              bit_rate_prob=arr("bp", bit_rate_prob if nbr else [1.0], np.float64),
              link_order=arr("lo", t.link_order, np.int32))
    return cfg, tab, keep


class OracleEnv:
    """One reference-semantics environment (gym 0.21 API shape: reset/step/observation)."""

    def __init__(self, env_id, tables, **kw):
        self.L = lib()
        self.cfg, self.tab, self._keep = make_cfg_tab(env_id, tables, **kw)
        self.kind = self.cfg.kind
        self.adim = ACTION_DIM[self.kind]
        self.h = self.L.oracle_create(C.byref(self.cfg), C.byref(self.tab))
        n = tables.num_nodes
        self.obs_dim = 1 + 2 * n + (2 * self.cfg.j + 3) * tables.k_paths
        self.cells = (self.cfg.num_cores, self.cfg.num_links, self.cfg.num_slots)

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    __del__ = close

    def set_trace(self, arrival, holding, src, dst, bit_rate=None):
        self._trace = [np.ascontiguousarray(arrival, np.float64), np.ascontiguousarray(holding, np.float64),
                       np.ascontiguousarray(src, np.int32), np.ascontiguousarray(dst, np.int32),
                       None if bit_rate is None else np.ascontiguousarray(bit_rate, np.int32)]
        self.L.oracle_set_trace(self.h, *[_ptr(a) for a in self._trace], len(self._trace[0]))

    def set_philox(self, seed, env_index):
        self.L.oracle_set_philox(self.h, int(seed), int(env_index))

    def reset(self, full=False):
        self.L.oracle_reset(self.h, int(full))

    def step(self, action):
        a = np.zeros(4, np.int32)
        a[:self.adim] = np.atleast_1d(np.asarray(action, np.int64)).astype(np.int32)[:self.adim]
        o = StepOut()
        rc = self.L.oracle_step(self.h, _ptr(a), C.byref(o))
        return o, rc

    def observation(self):
        obs = np.zeros(self.obs_dim, np.float64)
        self.L.oracle_observation(self.h, _ptr(obs))
        return obs

    def observation_int(self):
        out = np.zeros((self.cfg.k_paths, 2 * self.cfg.j + 3), np.int32)
        self.L.oracle_observation_int(self.h, _ptr(out))
        return out

    def heuristic(self, which):
        a = np.zeros(4, np.int32)
        self.L.oracle_heuristic(self.h, int(which), _ptr(a))
        return a[:self.adim].copy()

    def random_action(self):
        a = np.zeros(4, np.int32)
        self.L.oracle_random_action(self.h, _ptr(a))
        return a[:self.adim].copy()

    def request(self):
        arr, hold, ints = C.c_double(), C.c_double(), np.zeros(4, np.int32)
        self.L.oracle_get_request(self.h, C.byref(arr), C.byref(hold), _ptr(ints))
        return dict(arrival=arr.value, holding=hold.value, src=int(ints[0]), dst=int(ints[1]),
                    bit_rate=int(ints[2]), service_id=int(ints[3]))

    def counters(self):
        c = np.zeros(8, np.int64)
        self.L.oracle_get_counters(self.h, _ptr(c))
        return c

    def bit_rate_blocking(self):
        """info["bit_rate_blocking_<rate>"] per discrete bit rate + info["fairness"] of the last step."""
        out = np.zeros(self.cfg.num_bit_rates + 1, np.float64)
        self.L.oracle_get_bit_rate_blocking(self.h, _ptr(out))
        return out

    def action_probability(self):
        """RWA: (info["path_action_probability"], info["wavelength_action_probability"]) after the last step."""
        rej = int(self.cfg.allow_rejection)
        R, Cn = self.cfg.k_paths + rej, self.cfg.num_slots + rej
        out = np.zeros(R + Cn, np.float64)
        self.L.oracle_get_action_probability(self.h, _ptr(out))
        return out[:R].copy(), out[R:].copy()

    def link_stats(self):
        """Row f1 graph attributes: (per link [E, 3] = utilization, external_fragmentation, compactness; [2] = graph throughput,
        compactness) as of now."""
        link = np.zeros((self.cfg.num_links, 3), np.float64)
        graph = np.zeros(2, np.float64)
        self.L.oracle_get_link_stats(self.h, _ptr(link), _ptr(graph))
        return link, graph

    def state(self):
        avail = np.zeros(self.cells, np.int8)
        alloc = np.zeros(self.cells, np.int32)
        now, nheap = C.c_double(), C.c_int32()
        self.L.oracle_get_state(self.h, _ptr(avail), _ptr(alloc), C.byref(now), C.byref(nheap))
        return avail, alloc, now.value, nheap.value

    def error(self):
        return self.L.oracle_error(self.h)

    def rollout(self, T, policy=1, actions=None, want_obs=False):
        """policy 0: given actions [T,adim]; 1: Philox-random; 10+h: heuristic h.  Auto-resets on done."""
        dec = np.zeros((T, 4), np.int32)
        rew = np.zeros(T, np.float64)
        done = np.zeros(T, np.uint8)
        obs = np.zeros((T, self.obs_dim), np.float64) if want_obs else None
        aout = np.zeros((T, self.adim), np.int32)
        if actions is not None:
            actions = np.ascontiguousarray(actions, np.int32).reshape(T, self.adim)
        self.L.oracle_rollout(self.h, T, policy, _ptr(actions), self.adim, _ptr(dec), _ptr(rew), _ptr(done),
                              _ptr(obs), self.obs_dim, _ptr(aout))
        return dict(decisions=dec, rewards=rew, dones=done, obs=obs, actions=aout)


# ---------------------------------------------------------------- row f2: the reference's gym wrappers (numpy restatement)
def simple_matrix_observation(env, num_nodes):
    """SimpleMatrixObservation.observation (rmsa_env.py:806-837, rmcsa_env.py:914-947): float64 0/1 vector
    one-hot(min(src_id, dst_id)) ++ one-hot(max(src_id, dst_id)) ++ available_slots.reshape(-1)."""
    avail, _, _, _ = env.state()
    r = env.request()
    tau = np.zeros((2, num_nodes))
    tau[0, min(r["src"], r["dst"])] = 1
    tau[1, max(r["src"], r["dst"])] = 1
    spectrum = avail.astype(np.float64)
    if env.cfg.num_cores == 1:
        spectrum = spectrum[0]
    return np.concatenate((tau.reshape((1, tau.size)), spectrum.reshape((1, spectrum.size))), axis=1).reshape(-1)


def path_only_first_fit(env, tables, action):
    """PathOnlyFirstFitAction.action (rmsa_env.py:840-874; rwa_env.py:505-536): path index -> (path, first-fit
    slot); RMSA tries range(0, S - n) only, RWA range(W); otherwise the reject action (k, S)."""
    import math

    k, S = tables.k_paths, env.cfg.num_slots
    if action < k:
        avail, _, _, _ = env.state()
        r = env.request()
        row = tables.rows_of(r["src"], r["dst"])[action]
        links = tables.links_of(row)
        if env.kind == KINDS["RWA-v0"]:
            for w in range(S):
                if np.all(avail[0, links, w] == 1):
                    return (action, w)
        else:
            n = math.ceil(r["bit_rate"] / (int(tables.path_se[row]) * env.cfg.channel_width)) + 1
            for s in range(0, S - n):
                if np.all(avail[0, links, s:s + n] == 1):
                    return (action, s)
    return (k, S)


def rollout_mt(env_id, tables, *, seed, env0, n_envs, steps, policy=1, with_obs=True, threads=None, **kw):
    """All-host-threads throughput run (bench.py cpu_baseline / --impl reference). Returns accepted count."""
    cfg, tab, keep = make_cfg_tab(env_id, tables, **kw)
    threads = threads or os.cpu_count() or 1
    return lib().oracle_rollout_mt(C.byref(cfg), C.byref(tab), int(seed), int(env0), int(n_envs), int(steps),
                                   int(policy), int(with_obs), int(threads))


class OracleVec:
    """Persistent batch of oracle envs stepped by all host threads: the CPU arm of bench.py
    (the reference's one-env-per-worker SubprocVecEnv pattern, restated in C)."""

    def __init__(self, env_id, tables, n_envs, *, seed, env0=0, threads=None, **kw):
        self.L = lib()
        self.cfg, self.tab, self._keep = make_cfg_tab(env_id, tables, **kw)
        self.n_envs = int(n_envs)
        self.threads = int(threads or os.cpu_count() or 1)
        self.h = self.L.oracle_vec_create(C.byref(self.cfg), C.byref(self.tab), int(seed), int(env0), self.n_envs)

    def run(self, steps, policy=1, with_obs=True):
        """Advance every env by `steps` steps; returns the number of accepted requests."""
        return self.L.oracle_vec_run(self.h, int(steps), int(policy), int(with_obs), self.threads)

    def counters(self, i):
        c = np.zeros(8, np.int64)
        self.L.oracle_get_counters(self.L.oracle_vec_env(self.h, int(i)), _ptr(c))
        return c

    def state(self, i):
        """(available_slots int8 [C,E,S], spectrum_slots_allocation int32 [C,E,S], current_time, len(_events)) of env i."""
        cells = (self.cfg.num_cores, self.cfg.num_links, self.cfg.num_slots)
        avail, alloc = np.zeros(cells, np.int8), np.zeros(cells, np.int32)
        now, nheap = C.c_double(), C.c_int32()
        self.L.oracle_get_state(self.L.oracle_vec_env(self.h, int(i)), _ptr(avail), _ptr(alloc), C.byref(now), C.byref(nheap))
        return avail, alloc, now.value, nheap.value

    def close(self):
        if self.h:
            self.L.oracle_vec_destroy(self.h)
            self.h = None

    __del__ = close
