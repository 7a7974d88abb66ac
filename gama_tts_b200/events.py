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


def synthetic_events(seed, n_postures, special_rate=0.02, tight=False, grid=4, duration_ms=None):
    """An event list of about 4 events per posture: times ascending on a grid of `grid` ms, 2 to 15 grid steps apart, as the
    rule engine's are (with `tight`: any millisecond, some closer than a period or at equal times -- the reference then
    divides by zero and so do we), with `duration_ms` (a multiple of `grid`) the list is cut so that its last event falls exactly
    there (duration_ms / control period frames when the period is the grid: batches of one length), the first event with every
    parameter set, later ones with about half of them, a few
    special parameters, a macro-intonation polynomial on about one event in ten, the last event with every parameter."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = max(2, 4 * n_postures)
    ev = np.zeros(n, capi.EVENT_DTYPE)
    if tight:
        gaps = rng.integers(1, 60, n)
        gaps[rng.random(n) < 0.15] = 0
    else:
        gaps = grid * rng.integers(2, 16, n)
    gaps[0] = 0
    times = np.cumsum(gaps)
    if duration_ms is not None:
        assert not tight and duration_ms % grid == 0
        while times[-1] < duration_ms:                       # not enough postures for the duration: more events
            more = grid * rng.integers(2, 16, n)
            times = np.concatenate([times, times[-1] + np.cumsum(more)])
        n = int(np.searchsorted(times, duration_ms - grid, side="right")) + 1      # events before the end, then the last one
        times = np.concatenate([times[:n - 1], [duration_ms]])
        ev = np.zeros(n, capi.EVENT_DTYPE)
    ev["time"] = times
    targets = tracks.POSTURE_TARGETS[rng.integers(0, len(tracks.POSTURE_TARGETS), n)].astype(np.float64)
    targets += rng.normal(0.0, 0.01, targets.shape)
    span = tracks.PARAM_MAX - tracks.PARAM_MIN
    targets = np.clip(targets, tracks.PARAM_MIN + 0.02 * span, tracks.PARAM_MAX - 0.02 * span)   # room for the special parameters
    keep = rng.random((n, 16)) < 0.5
    keep[0] = keep[-1] = True
    ev["param"] = np.where(keep, targets, EMPTY)
    ev["param"][1:, 0] = np.where(keep[1:, 0], rng.normal(0.0, 1.0, n - 1), EMPTY)       # micro intonation
    sp = rng.random((n, 16)) < special_rate
    ev["special"] = np.where(sp, np.clip(rng.normal(0.0, 0.003, (n, 16)), -0.01, 0.01) * span, EMPTY)
    # macro intonation: from an event on, y0 + m (x - t0) + q (x - t0)^3 semitones in the chunk's time x (ms), as
    # coefficients of x (the reference fits them through its intonation points: a few semitones over a tone group)
    interp = rng.random(n) < 0.1
    interp[0] = tight and interp[0]       # a polynomial on the event at time 0 makes the first segment's slope x / 0
    t0 = ev["time"].astype(np.float64)
    y0, m, q = rng.normal(-2.0, 2.0, n), rng.uniform(-0.002, 0.002, n), rng.uniform(-2e-10, 2e-10, n)
    ev["has_interp"] = interp
    ev["a"] = np.where(interp, q, 0.0)
    ev["b"] = np.where(interp, -3.0 * q * t0, 0.0)
    ev["c"] = np.where(interp, m + 3.0 * q * t0 * t0, 0.0)
    ev["d"] = np.where(interp, y0 - m * t0 - q * t0 ** 3, 0.0)
    return ev
