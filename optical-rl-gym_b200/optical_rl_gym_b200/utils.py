"""Batched counterparts of ``optical_rl_gym/utils.py``: ``evaluate_heuristic`` (utils.py:103-141) and
``random_policy`` (utils.py:98-99) for an :class:`OpticalVecEnv`."""
from __future__ import annotations

import torch


def random_policy(env):
    """``env.action_space.sample()`` for every env (utils.py:98-99): the device-side uniform policy."""
    return env.sample_actions()


def evaluate_heuristic(env, heuristic="shortest_available_path_first_fit", n_eval_episodes: int = 10, render: bool = False,
                       callback=None, reward_threshold=None, return_episode_rewards: bool = False, chunk: int = 256):
    """``utils.evaluate_heuristic`` for a batch: every env plays ``n_eval_episodes`` episodes (``env.reset()`` before each,
    then ``action = heuristic(env); env.step(action)`` until ``done``), as rollouts of ``chunk`` steps on the device
    (``OpticalVecEnv.rollout``: one persistent kernel launch per chunk where it applies).

    ``heuristic``: a heuristic name (``"shortest_available_path_first_fit"``, ``"sap_ff"`` ...), ``"random"`` /
    :func:`random_policy`.  Returns ``(mean_reward, std_reward)`` over all ``num_envs * n_eval_episodes`` episodes (the
    reference's ``np.mean`` / ``np.std`` over its episode list), or ``(episode_rewards [episodes, N], episode_lengths
    [episodes, N])`` tensors with ``return_episode_rewards``.  ``render`` / ``callback`` are not supported (no per-step host
    round trip)."""
    if render or callback is not None:
        raise NotImplementedError("render / callback would need a host round trip per step")
    if heuristic is random_policy or heuristic is None:
        heuristic = "random"
    if callable(heuristic):
        heuristic = getattr(heuristic, "__name__", str(heuristic))
    n = env.num_envs
    # every env of a handle is reset and stepped together, so the episodes end at the same steps
    steps_per_episode = env.episode_length if env.env_id == "RWA-v0" else env.episode_length - 1
    env.reset()
    total = n_eval_episodes * steps_per_episode
    ep_rewards = torch.zeros((n_eval_episodes, n), dtype=torch.float64, device=env.device)
    ep_lengths = torch.zeros((n_eval_episodes, n), dtype=torch.int64, device=env.device)
    running = torch.zeros(n, dtype=torch.float64, device=env.device)
    length, episode, t0 = 0, 0, 0
    while t0 < total:
        t = min(chunk, total - t0)
        _, reward, done, _ = env.rollout(t, heuristic, want_obs=False, want_actions=False)
        ends = done.any(dim=1)
        if not torch.equal(ends, done.all(dim=1)):
            raise RuntimeError("episodes of this handle are not in lock step")
        csum = torch.cumsum(reward.to(torch.float64), dim=0)
        prev = torch.zeros(n, dtype=torch.float64, device=env.device)
        last = -1
        for b in torch.nonzero(ends).flatten().tolist():
            ep_rewards[episode] = running + csum[b] - prev
            ep_lengths[episode] = length + (b - last)
            running.zero_()
            prev, last, length = csum[b].clone(), b, 0
            episode += 1
        running += csum[-1] - prev
        length += (t - 1 - last)
        t0 += t
    assert episode == n_eval_episodes, (episode, n_eval_episodes)
    if return_episode_rewards:
        return ep_rewards, ep_lengths
    mean_reward, std_reward = ep_rewards.mean().item(), ep_rewards.std(unbiased=False).item()
    if reward_threshold is not None:
        assert mean_reward > reward_threshold, "Mean reward below threshold: {:.2f} < {:.2f}".format(mean_reward, reward_threshold)
    return mean_reward, std_reward
