"""Event lists for the control-frame generation path (gtts_events_*): the configuration record of a chunk and synthetic
event lists shaped like the ones the reference's rule engine builds (vtm_control_model/EventList.cpp:generateEventList,
applyIntonation) -- for the tests and the benchmark, which have no text front end on the GPU box."""
import numpy as np

from . import capi, tracks

EMPTY = np.inf      # Event::EMPTY_PARAMETER


def event_config(control_period=4, macro=1, micro=1, drift=1, smooth=1, initial_pitch=-20.0, mean_pitch=-16.0,
                 drift_deviation=4.0, drift_lowpass_cutoff=4.0):
    """The record of a chunk with a fresh drift generator; the defaults are data/voice/english/0_male/vtm_control_model.txt's."""
    cfg = np.zeros(1, capi.EVENT_CONFIG_DTYPE)
    cfg["control_period"], cfg["macro_intonation"], cfg["micro_intonation"] = control_period, macro, micro
    cfg["intonation_drift"], cfg["smooth_intonation"] = drift, smooth
    cfg["initial_pitch"], cfg["mean_pitch"] = initial_pitch, mean_pitch
    capi.check(capi.load().gtts_events_drift_setup(float(drift_deviation), 1000.0 / control_period, float(drift_lowpass_cutoff),
                                                   cfg.ctypes.data))
    return cfg[0]


def synthetic_events(seed, n_postures, special_rate=0.02, tight=False):
    """An event list of about 4 events per posture: times ascending (not multiples of the control period; with `tight` some
    closer than a period or equal), the first event with every parameter set, later ones with about half of them, a few
    special parameters, a macro-intonation polynomial on about one event in ten, the last event with every parameter."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = max(2, 4 * n_postures)
    ev = np.zeros(n, capi.EVENT_DTYPE)
    gaps = rng.integers(1 if tight else 5, 60, n)
    if tight:
        gaps[rng.random(n) < 0.15] = 0
    gaps[0] = 0
    ev["time"] = np.cumsum(gaps)
    targets = tracks.POSTURE_TARGETS[rng.integers(0, len(tracks.POSTURE_TARGETS), n)].astype(np.float64)
    targets += rng.normal(0.0, 0.01, targets.shape)
    keep = rng.random((n, 16)) < 0.5
    keep[0] = keep[-1] = True
    ev["param"] = np.where(keep, targets, EMPTY)
    ev["param"][1:, 0] = np.where(keep[1:, 0], rng.normal(0.0, 1.0, n - 1), EMPTY)       # micro intonation
    sp = rng.random((n, 16)) < special_rate
    ev["special"] = np.where(sp, rng.normal(0.0, 0.05, (n, 16)), EMPTY)
    interp = rng.random(n) < 0.1
    ev["has_interp"] = interp
    ev["a"] = np.where(interp, rng.normal(0.0, 1e-8, n), 0.0)
    ev["b"] = np.where(interp, rng.normal(0.0, 1e-5, n), 0.0)
    ev["c"] = np.where(interp, rng.normal(0.0, 1e-2, n), 0.0)
    ev["d"] = np.where(interp, rng.normal(-2.0, 2.0, n), 0.0)
    return ev
