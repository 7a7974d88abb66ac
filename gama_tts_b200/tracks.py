"""Synthetic control-parameter tracks for tests and benchmarks.

A track is what ``EventList::generateOutput`` hands to ``Controller::synthesize`` in the reference
(gama_tts/src/vtm_control_model/EventList.cpp:929-1091, Controller.cpp:277-313): F control frames of 16
float32 parameters at the control rate (250 Hz), order as in ``artic.xml:38-53``:
pitch(microInt), glotVol, aspVol, fricVol, fricPos, fricCF, fricBW, r1..r8, velum.

The generator walks randomly over the posture targets of the reference's articulatory database
(``data/voice/english/0_male/artic.xml`` -- the 65 ``<posture>`` parameter-target rows are embedded below
as data) with linear ramps and holds on the 4 ms grid, plus a smooth pitch random walk, so that the
fraction of frames in which each parameter changes is close to that of real front-end tracks
(pitch 100 percent, radii 55-88, glotVol 20-67, frication 1-50).  Seeded with PCG64: a given
(seed, n_frames) always yields the same float32 track.
"""
import numpy as np

PARAM_NAMES = ("pitch", "glotVol", "aspVol", "fricVol", "fricPos", "fricCF", "fricBW",
               "r1", "r2", "r3", "r4", "r5", "r6", "r7", "r8", "velum")
PARAM_MIN = np.array([-10, 0, 0, 0, 0, 100, 250, 0, 0, 0, 0, 0, 0, 0, 0, 0], np.float64)
PARAM_MAX = np.array([10, 60, 60, 10, 7, 20000, 20000, 3, 3, 3, 3, 3, 3, 3, 3, 1.5], np.float64)

# (symbol, microInt, glotVol, aspVol, fricVol, fricPos, fricCF, fricBW, r1, ..., r8, velum)
POSTURES = (
    ('#', 0, 0, 0, 0, 5.5, 2500, 500, 0.8, 0.89, 0.99, 0.81, 0.76, 1.05, 1.23, 0.01, 0.1),
    ('^', 0, 0, 0, 0, 5.5, 2500, 500, 0.8, 0.89, 0.99, 0.81, 0.76, 1.05, 1.23, 0.01, 0.1),
    ('a', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.65, 0.65, 0.65, 1.31, 1.23, 1.31, 1.67, 0.1),
    ('aa', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.65, 0.84, 1.15, 1.31, 1.59, 1.59, 2.61, 0.1),
    ('ah', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.65, 0.45, 0.94, 1.1, 1.52, 1.46, 2.45, 0.1),
    ('an', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.52, 0.45, 0.79, 1.49, 1.67, 1.02, 1.59, 1.5),
    ('ar', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.52, 0.45, 0.79, 1.49, 1.67, 1.02, 1.59, 0.1),
    ('aw', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 1.1, 0.94, 0.42, 1.49, 1.67, 1.78, 1.05, 0.1),
    ('b', -2, 43.5, 0, 0, 7, 2000, 700, 0.8, 0.89, 0.76, 1.28, 1.8, 0.99, 0.84, 0.1, 0.1),
    ('bx', -2, 43.5, 0, 0, 7, 2000, 700, 0.8, 0.89, 0.76, 1.28, 1.8, 0.99, 0.84, 0.1, 0.1),
    ('ch', -2, 0, 0, 0, 5.6, 2500, 2600, 0.8, 1.36, 1.74, 1.87, 0.94, 0, 0.79, 0.79, 0.1),
    ('d', -2, 43.5, 0, 0, 6.7, 4500, 2000, 0.8, 1.31, 1.49, 1.25, 0.76, 0.1, 1.44, 1.3, 0.1),
    ('dh', -1, 54, 0, 0.25, 6, 4400, 4500, 0.8, 1.2, 1.5, 1.35, 1.2, 1.2, 0.4, 1, 0.1),
    ('dx', -2, 43.5, 0, 0, 6.7, 4500, 2000, 0.8, 1.31, 1.49, 1.25, 0.76, 0.1, 1.44, 1.31, 0.1),
    ('e', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.68, 1.12, 1.695, 1.385, 1.07, 1.045, 2.06, 0.1),
    ('ee', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 1.67, 1.905, 1.985, 0.81, 0.495, 0.73, 1.485, 0.1),
    ('er', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.885, 0.99, 0.81, 0.755, 1.045, 1.225, 1.12, 0.1),
    ('f', -1, 0, 0, 0.5, 7, 3300, 1000, 0.8, 0.89, 0.99, 0.81, 0.76, 0.89, 0.84, 0.5, 0.1),
    ('g', -2, 43.5, 0, 0, 4.7, 2000, 2000, 0.8, 1.7, 1.3, 0.99, 0.1, 1.07, 0.73, 1.49, 0.1),
    ('gs', 0, 0, 0, 0, 5.5, 2500, 500, 0.8, 0.8, 0.8, 0.8, 0.8, 0.8, 0.8, 0.8, 0.1),
    ('h', 0, 0, 10, 0, 5.5, 2500, 500, 0.8, 0.8, 0.8, 0.8, 0.8, 0.8, 0.8, 0.8, 0.1),
    ('hh', 0, 0, 10, 0, 1, 1000, 1000, 0.8, 0.24, 0.4, 0.81, 0.76, 1.05, 1.23, 1.12, 0.1),
    ('hv', 0, 42, 10, 0, 5.5, 2500, 500, 0.8, 0.8, 0.8, 0.8, 0.8, 0.8, 0.8, 0.8, 0.1),
    ('i', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 1.045, 1.565, 1.75, 0.94, 0.68, 0.785, 1.12, 0.1),
    ('in', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.65, 0.835, 1.15, 1.305, 1.59, 1.59, 2.61, 1.5),
    ('j', -2, 48, 0, 0, 5.6, 2500, 2600, 0.8, 1.36, 1.74, 1.87, 0.94, 0, 0.79, 0.79, 0.1),
    ('k', -10, 0, 0, 0, 4.7, 2000, 2000, 0.8, 1.7, 1.3, 0.99, 0.1, 1.07, 0.73, 1.49, 0.1),
    ('kx', -10, 0, 0, 0, 4.7, 2000, 2000, 0.8, 1.7, 1.3, 0.99, 0.1, 1.07, 0.73, 1.49, 0.1),
    ('l', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.89, 1.1, 0.97, 0.89, 0.34, 0.29, 1.12, 0.1),
    ('ll', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.63, 0.47, 0.65, 1.54, 0.45, 0.26, 1.05, 0.1),
    ('ls', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.63, 0.47, 0.65, 1.54, 0.45, 0.26, 1.05, 0.1),
    ('m', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.89, 0.76, 1.28, 1.8, 0.99, 0.84, 0.1, 0.5),
    ('n', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 1.31, 1.49, 1.25, 1, 0.05, 1.44, 1.31, 0.5),
    ('ng', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 1.7, 1.3, 0.99, 0.1, 1.07, 0.73, 1.49, 0.5),
    ('o', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 1, 0.925, 0.6, 1.27, 1.83, 1.97, 1.12, 0.1),
    ('oh', 0, 0, 0, 0, 5.5, 2500, 500, 0.8, 0.885, 0.99, 0.81, 0.755, 1.045, 1.225, 1.12, 0.1),
    ('on', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 1, 0.925, 0.6, 1.265, 1.83, 1.965, 1.12, 1.5),
    ('ov', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.885, 0.99, 0.81, 0.755, 1.045, 1.225, 1.12, 0.1),
    ('p', -10, 0, 0, 0, 7, 2000, 700, 0.8, 0.89, 0.76, 1.28, 1.8, 0.99, 0.84, 0.1, 0.1),
    ('ph', -1, 0, 0, 24, 7, 864, 3587, 0.8, 0.89, 0.99, 0.81, 0.6, 0.52, 0.71, 0.24, 0.1),
    ('px', -10, 0, 0, 0, 7, 2000, 700, 0.8, 0.89, 0.76, 1.28, 1.8, 0.99, 0.84, 0.1, 0.1),
    ('q', 0, 0, 0, 0, 5.5, 2500, 500, 0.8, 0.89, 0.99, 0.81, 0.76, 1.05, 1.23, 0.01, 0.1),
    ('qc', -2, 0, 0, 0, 5.6, 2500, 2600, 0.8, 1.36, 1.74, 1.87, 0.94, 0.1, 0.79, 0.79, 0.1),
    ('qk', -10, 0, 0, 0, 4.7, 2000, 2000, 0.8, 1.7, 1.3, 0.99, 0.1, 1.07, 0.73, 1.49, 0.1),
    ('qp', -10, 0, 0, 0, 7, 2000, 700, 0.8, 0.89, 0.76, 1.28, 1.8, 0.99, 0.84, 0.1, 0.1),
    ('qs', 0, 0, 0, 0, 5.8, 5500, 500, 0.8, 1.31, 1.49, 1.25, 0.9, 0.2, 0.4, 1.31, 0.1),
    ('qt', -10, 0, 0, 0, 7, 4500, 2000, 0.8, 1.31, 1.49, 1.25, 0.76, 0.1, 1.44, 1.31, 0.1),
    ('qz', -1, 0, 0, 0, 5.8, 5500, 500, 0.8, 1.31, 1.49, 1.25, 0.9, 0.2, 0.6, 1.31, 0.1),
    ('r', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 1.31, 0.73, 1.07, 2.12, 0.47, 1.78, 0.65, 0.1),
    ('rr', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 1.31, 0.73, 1.31, 2.12, 0.63, 1.78, 0.65, 0.1),
    ('s', 0, 0, 0, 0.8, 5.8, 5500, 500, 0.8, 1.31, 1.49, 1.25, 0.9, 0.2, 0.4, 1.31, 0.1),
    ('sh', 0, 0, 0, 0.4, 5.6, 2500, 2600, 0.8, 1.36, 1.74, 1.87, 0.94, 0.37, 0.79, 0.79, 0.1),
    ('t', -10, 0, 0, 0, 7, 4500, 2000, 0.8, 1.31, 1.49, 1.25, 0.76, 0.1, 1.44, 1.31, 0.1),
    ('th', 0, 0, 0, 0.25, 6, 4400, 4500, 0.8, 1.2, 1.5, 1.35, 1.2, 1.2, 0.4, 1, 0.1),
    ('tx', -10, 0, 0, 0, 6.7, 4500, 2000, 0.8, 1.31, 1.49, 1.25, 0.76, 0.1, 1.44, 1.31, 0.1),
    ('u', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.625, 0.6, 0.705, 1.12, 1.93, 1.515, 0.625, 0.1),
    ('uh', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.89, 0.99, 0.81, 0.76, 1.05, 1.23, 1.12, 0.1),
    ('un', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.885, 0.99, 0.81, 0.755, 1.045, 1.225, 1.12, 1.5),
    ('uu', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 1.91, 1.44, 0.6, 1.02, 1.33, 1.56, 0.55, 0.1),
    ('v', -1, 54, 0, 0.2, 7, 3300, 1000, 0.8, 0.89, 0.99, 0.81, 0.76, 0.89, 0.84, 0.5, 0.1),
    ('w', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 1.91, 1.44, 0.6, 1.02, 1.33, 1.56, 0.55, 0.1),
    ('x', 0, 0, 0, 0.5, 2, 1770, 900, 0.8, 1.7, 1.3, 0.4, 0.99, 1.07, 0.73, 1.49, 0.1),
    ('y', 0, 60, 0, 0, 5.5, 2500, 500, 0.8, 1.67, 1.91, 1.99, 0.63, 0.29, 0.58, 1.49, 0.25),
    ('z', -1, 54, 0, 0.8, 5.8, 5500, 500, 0.8, 1.31, 1.49, 1.25, 0.9, 0.2, 0.6, 1.31, 0.1),
    ('zh', -1, 54, 0, 0.4, 5.6, 2500, 2600, 0.8, 1.36, 1.74, 1.87, 0.94, 0.37, 0.79, 0.79, 0.1),
)

POSTURE_TARGETS = np.array([p[1:] for p in POSTURES], np.float64)

VOWEL_AA = np.array([-12, 60, 0, 0, 5.5, 2500, 500, 0.8, 0.65, 0.84, 1.15, 1.31, 1.59, 1.59, 2.61, 0.1], np.float32)


def synthetic_track(seed, n_frames, pitch_center=-12.0, pitch_lo=-24.0, pitch_hi=0.0):
    """[n_frames, 16] float32: random posture walk + pitch random walk (BASELINE configs 2 and 3)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_frames = int(n_frames)
    out = np.empty((n_frames, 16), np.float64)
    cur = POSTURE_TARGETS[0].copy()          # start from silence ('#')
    f = 0
    while f < n_frames:
        tgt = POSTURE_TARGETS[int(rng.integers(0, len(POSTURE_TARGETS)))]
        ramp = int(rng.integers(10, 31))     # 40-120 ms transition
        hold = int(rng.integers(5, 21))      # 20-80 ms quasi-steady state
        n = min(ramp, n_frames - f)
        w = (np.arange(1, n + 1) / ramp)[:, None]
        out[f:f + n] = cur + (tgt - cur) * w
        f += n
        cur = tgt.copy()
        n = min(hold, n_frames - f)
        out[f:f + n] = cur
        f += n
    out = np.clip(out, PARAM_MIN, PARAM_MAX)
    # pitch: smooth random walk in semitones (reflected at the limits), replaces the microInt column
    walk = pitch_center + np.cumsum(rng.normal(0.0, 0.12, n_frames))
    span = pitch_hi - pitch_lo
    folded = np.mod(walk - pitch_lo, 2.0 * span)
    out[:, 0] = pitch_lo + np.where(folded > span, 2.0 * span - folded, folded)
    return out.astype(np.float32)


def config2_tracks(n_utt=1024, n_frames=2500, seed0=20240):
    """BASELINE config 2: n_utt tracks of n_frames frames (10 s), seeds seed0 + u."""
    return [synthetic_track(seed0 + u, n_frames) for u in range(n_utt)]


def config3_lengths(n_utt=65536, seed=7, lo=250, hi=5000):
    """BASELINE config 3 lengths: log-uniform in [lo, hi] frames."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.exp(rng.uniform(np.log(lo), np.log(hi), n_utt)).astype(np.int64)


def config3_utterance(u, lengths=None):
    """Utterance u of THE BASELINE config 3 draw: length from the log-uniform draw of
    config3_lengths(65536, seed 7), its own randomised voice from PCG64(7 + u), its own track (seed 7 + u).
    Returns (voice dict, track [F, 16] float32)."""
    from .voices import random_voice
    if lengths is None:
        lengths = config3_lengths()
    voice = random_voice(np.random.Generator(np.random.PCG64(7 + int(u))))
    return voice, synthetic_track(7 + int(u), int(lengths[u]))


def tile_track(track, n_frames):
    """Repeats a track to n_frames frames (used to build long tracks from a short seed track)."""
    reps = -(-int(n_frames) // len(track))
    return np.ascontiguousarray(np.tile(track, (reps, 1))[:n_frames], np.float32)
