"""Times the UNMODIFIED reference (optical_rl_gym from baseline/_ref, installed by __graft_entry__.build() with
`pip install --no-deps --target baseline/_ref`) on the host cores: one env per worker process, driven step by step
through pipes exactly like Stable-Baselines3's SubprocVecEnv drives it (SB3 itself is not installable here, so the
~40-line protocol is restated below).  gym / matplotlib come from the import shim in tests/golden/refshim.

Used by bench.py (`cpu_baseline_python`); a reported baseline, not a target."""
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def available():
    return os.path.isdir(os.path.join(REF_DIR, "optical_rl_gym")) and \
        os.path.exists(os.path.join(REF_DIR, "examples", "topologies", "nsfnet_chen_5-paths_6-modulations.h5"))


def _worker(conn, env_id, env_args, seed):
    os.environ["ORLG_REFERENCE"] = REF_DIR
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_harness as rh

    topo = rh.load_topology()
    env = rh.make(env_id, topology=topo, seed=seed, **env_args)
    obs = env.reset()
    conn.send(("ready", obs))
    while True:
        cmd, data = conn.recv()
        if cmd == "step":
            obs, reward, done, info = env.step(data)
            if done:
                info["terminal_observation"] = obs
                obs = env.reset()
            conn.send((obs, reward, done, info))
        elif cmd == "close":
            conn.close()
            return


class SubprocVecEnv:
    """VecEnv over worker processes: step_async sends one action per pipe, step_wait gathers (obs, reward, done, info)."""

    def __init__(self, env_id, env_args, n_workers, seed0=1):
        ctx = mp.get_context("fork")
        self.conns, self.procs = [], []
        for i in range(n_workers):
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(child, env_id, env_args, seed0 + i), daemon=True)
            p.start()
            child.close()
            self.conns.append(parent); self.procs.append(p)
        self.obs = [c.recv()[1] for c in self.conns]

    def step(self, actions):
        for c, a in zip(self.conns, actions):
            c.send(("step", a))
        return [c.recv() for c in self.conns]

    def close(self):
        for c in self.conns:
            try:
                c.send(("close", None))
            except Exception:  # noqa: BLE001
                pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()


def measure(env_id="DeepRMSA-v0", env_args=None, workers=None, warm_steps=1000, seconds=8.0, n_actions=5):
    """Aggregate env-steps/s of `workers` reference envs stepped in lock step with uniform random actions."""
    import numpy as np

    workers = workers or os.cpu_count() or 1
    env_args = dict(env_args or {})
    vec = SubprocVecEnv(env_id, env_args, workers)
    rng = np.random.default_rng(0)
    try:
        for _ in range(warm_steps):
            vec.step(rng.integers(0, n_actions, workers).tolist())
        steps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            vec.step(rng.integers(0, n_actions, workers).tolist())
            steps += 1
        dt = time.perf_counter() - t0
    finally:
        vec.close()
    return {"value": workers * steps / dt, "unit": "env-steps/s", "cores": workers, "kind": "reference",
            "sample": "%d reference envs (one per process, pipe-driven SubprocVecEnv protocol), %d lock steps in %.1f s "
                      "after a %d-step warm-up, uniform random actions" % (workers, steps, dt, warm_steps)}


if __name__ == "__main__":
    import json

    print(json.dumps(measure(seconds=float(sys.argv[1]) if len(sys.argv) > 1 else 8.0)))
