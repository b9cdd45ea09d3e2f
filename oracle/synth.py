"""
TEST INFRASTRUCTURE (see oracle/__init__.py) -- seeded synthetic trial
generators following the reference's own (syncopy/synthdata/analog.py,
syncopy/synthdata/utils.py:53-55, syncopy/tests/helpers.py:18 `test_seed=42`).
These define the benchmark / parity inputs.
"""
import numpy as np

TEST_SEED = 42


def trial_seeds(n_trials, seed=TEST_SEED):
    """One seed per trial, exactly like `collect_trials` (synthdata/utils.py:53-55)."""
    return np.random.default_rng(seed).integers(1_000_000, size=n_trials)


def white_noise_trial(n_samples, n_channels, seed):
    """synthdata/analog.py:38-40."""
    return np.random.default_rng(seed).normal(size=(n_samples, n_channels)).astype("f4")


def white_noise(n_trials, n_samples, n_channels, seed=TEST_SEED):
    """[nTrials, nSamples, nChannels] float32 white noise, per-trial seeds."""
    return np.stack([white_noise_trial(n_samples, n_channels, s)
                     for s in trial_seeds(n_trials, seed)])


def ar2_network_trial(adj=None, n_samples=1000, alphas=(0.55, -0.8), seed=None):
    """synthdata/analog.py:185-252: coupled AR(2) processes, float32 state."""
    if adj is None:
        adj = np.zeros((2, 2), dtype=np.float32)
        adj[1, 0] = 0.25
    else:
        adj = adj.astype(np.float32)
    n_chan = adj.shape[0]
    a1, a2 = alphas
    step = np.diag(n_chan * [a1]) + adj.T
    sig = np.zeros((n_samples, n_chan), dtype=np.float32)
    rng = np.random.default_rng(seed)
    sig[:2] = rng.normal(size=(2, n_chan))
    for i in range(2, n_samples):
        sig[i] = step @ sig[i - 1] + a2 * sig[i - 2]
        sig[i] += rng.normal(size=n_chan)
    return sig


def ar2_network(n_trials, adj=None, n_samples=1000, alphas=(0.55, -0.8), seed=TEST_SEED):
    return np.stack([ar2_network_trial(adj, n_samples, alphas, s)
                     for s in trial_seeds(n_trials, seed)])


def harmonics_trial(freqs, amps, samplerate, n_samples, phases=None):
    """Sum-free multi-channel harmonics: channel c carries amps[c]*cos(2 pi freqs[c] t + phases[c])."""
    t = np.arange(n_samples) / samplerate
    freqs, amps = np.atleast_1d(freqs), np.atleast_1d(amps)
    if phases is None:
        phases = np.zeros(freqs.size)
    return (amps[None, :] * np.cos(2 * np.pi * freqs[None, :] * t[:, None] + phases[None, :])).astype("f4")
