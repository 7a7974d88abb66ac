"""Voice configurations: the key/value set VocalTractModel0 reads (reference
gama_tts/src/vtm/VocalTractModel0.h:266-305) from ``vtm.txt`` merged with ``variant/<name>.txt``
(gama_tts/src/vtm_control_model/Controller.cpp:48-49).

The numbers below are the shipped defaults of voice directory ``data/voice/english/0_male``
(``0_male/vtm.txt`` and ``0_male/variant/{male,female,large_child,small_child,baby}.txt``); they are
configuration data, not code.  ``load_voice_dir`` reads the same files from a voice directory with the
reference's parsing rules (ConfigurationData.cpp:67-118: ``key = value``, ``#`` comments at column 0,
duplicate keys are errors; Controller merges the variant file over vtm.txt).
"""
import os

import numpy as np

VTM_TXT = {
    "output_rate": 48000.0, "waveform": 0, "vocal_tract_length_offset": 0.0, "temperature": 32.0,
    "loss_factor": 0.8, "mouth_coefficient": 5000.0, "nose_coefficient": 5000.0, "throat_cutoff": 1500.0,
    "throat_volume": 6.0, "noise_modulation": 1, "mix_offset": 48.0,
}

_COMMON = {
    "aperture_radius": 3.05, "nasal_radius_1": 1.35, "nasal_radius_2": 1.96, "nasal_radius_3": 1.91,
    "nasal_radius_4": 1.3, "nasal_radius_5": 0.73, "global_nasal_radius_coef": 1.0, "global_radius_coef": 1.0,
    **{"radius_%d_coef" % i: 1.0 for i in range(1, 9)},
}

VARIANTS = {
    "male": {"vocal_tract_length": 17.5, "glottal_pulse_tp": 40.0, "glottal_pulse_tn_min": 24.0,
             "glottal_pulse_tn_max": 24.0, "breathiness": 0.5},
    "female": {"vocal_tract_length": 15.0, "glottal_pulse_tp": 40.0, "glottal_pulse_tn_min": 32.0,
               "glottal_pulse_tn_max": 32.0, "breathiness": 1.5},
    "large_child": {"vocal_tract_length": 12.5, "glottal_pulse_tp": 40.0, "glottal_pulse_tn_min": 24.0,
                    "glottal_pulse_tn_max": 24.0, "breathiness": 1.5},
    "small_child": {"vocal_tract_length": 10.0, "glottal_pulse_tp": 40.0, "glottal_pulse_tn_min": 24.0,
                    "glottal_pulse_tn_max": 24.0, "breathiness": 1.5},
    "baby": {"vocal_tract_length": 7.5, "glottal_pulse_tp": 40.0, "glottal_pulse_tn_min": 24.0,
             "glottal_pulse_tn_max": 24.0, "breathiness": 1.5},
}

# control_period = 4 ms in 0_male/vtm_control_model.txt -> controlRate = 1000 / 4
# (vtm_control_model/VTMControlModelConfiguration.cpp:37-41)
DEFAULT_CONTROL_RATE = 250.0

VOICE_KEYS = (list(VTM_TXT) + ["vocal_tract_length", "glottal_pulse_tp", "glottal_pulse_tn_min",
                               "glottal_pulse_tn_max", "breathiness"] + list(_COMMON))


def default_voice(variant="male"):
    """The merged configuration the reference's Controller hands to the tube model for 0_male/<variant>."""
    v = dict(VTM_TXT)
    v.update(_COMMON)
    v.update(VARIANTS[variant])
    return v


def parse_config_file(path):
    """key = value parser with the reference's rules (ConfigurationData.cpp:67-118)."""
    out = {}
    with open(path, "rb") as f:
        for ln, raw in enumerate(f.read().decode().split("\n"), 1):
            line = raw.rstrip("\r")
            if not line or line[0] == "#":
                continue
            if line[0].isspace():
                raise ValueError("%s:%d: space at the beginning of the line" % (path, ln))
            if "=" not in line:
                raise ValueError("%s:%d: missing separator" % (path, ln))
            key, value = line.split("=", 1)
            key, value = key.strip(), value.strip()
            if not key or not value:
                raise ValueError("%s:%d: empty key or value" % (path, ln))
            if key in out:
                raise ValueError("%s:%d: duplicate key %s" % (path, ln, key))
            out[key] = value
    return out


def load_voice_dir(voice_dir, variant=None):
    """Reads <voice_dir>/vtm.txt overlaid with variant/<variant>.txt, like Controller.cpp:40-55."""
    kv = parse_config_file(os.path.join(voice_dir, "vtm.txt"))
    if variant is None:
        cm = parse_config_file(os.path.join(voice_dir, "vtm_control_model.txt"))
        variant = cm["variant_name"]
    kv.update(parse_config_file(os.path.join(voice_dir, "variant", variant + ".txt")))
    voice = {}
    for k in VOICE_KEYS:
        voice[k] = int(kv[k]) if k in ("waveform", "noise_modulation") else float(kv[k])
    return voice


def random_voice(rng, base="male"):
    """Randomised voice of BASELINE config 3: tract length, glottal pulse shape,
    nasal coupling, breathiness.  ``rng`` is a numpy Generator."""
    v = default_voice(base)
    v["vocal_tract_length"] = float(rng.uniform(7.5, 17.5))
    v["glottal_pulse_tp"] = float(rng.uniform(30.0, 45.0))
    v["glottal_pulse_tn_min"] = float(rng.uniform(16.0, 28.0))
    v["glottal_pulse_tn_max"] = v["glottal_pulse_tn_min"] + float(rng.uniform(0.0, 12.0))
    for i in range(1, 6):
        v["nasal_radius_%d" % i] = _COMMON["nasal_radius_%d" % i] * float(rng.uniform(0.7, 1.3))
    v["breathiness"] = float(rng.uniform(0.5, 5.0))
    return v


def internal_rate(voice):
    """fs_int exactly as VocalTractModel0.h:343-344 (double arithmetic, truncation to int)."""
    length = voice["vocal_tract_length_offset"] + voice["vocal_tract_length"]
    length = min(max(length, 3.0), 30.0)
    c = 331.4 + (0.6 * voice["temperature"])
    return int((c * 10 * 100.0) / length)


def control_steps(voice, control_rate=DEFAULT_CONTROL_RATE):
    """Controller.cpp:286 (std::rint = round-half-even, like numpy)."""
    return int(np.rint(internal_rate(voice) / control_rate))


# ---- model 5 (VocalTractModel5.h:373-425): data/voice/english/5_male/vtm.txt + variant/<name>.txt ----------------------
VTM5_TXT = {"output_rate": 48000.0, "waveform": 0, "vocal_tract_length_offset": 0.0, "temperature": 35.0,
            "noise_modulation": 1, "bypass": 0}

_COMMON5 = {
    "glottal_pulse_tp": 40.0, "glottal_pulse_tn_min": 24.0, "glottal_pulse_tn_max": 24.0,
    "nasal_radius_2": 1.35, "nasal_radius_3": 1.83, "nasal_radius_4": 1.94, "nasal_radius_5": 1.72, "nasal_radius_6": 1.3,
    "nasal_radius_7": 0.73, **{"radius_%d_coef" % i: 1.0 for i in range(1, 9)},
    "loss_factor": 0.4, "mix_offset": 48.0, "glottal_noise_cutoff": 400.0, "frication_noise_cutoff": 5500.0,
    "frication_factor": 1.0, "min_glottal_loss": 0.0, "max_glottal_loss": 2.0, "glottal_lowpass_cutoff": 3000.0,
    "constant_radius_mouth_impedance": 0,
}

VARIANTS5 = {
    "male": {"vocal_tract_length": 17.5, "breathiness": 5.0, "global_nasal_radius_coef": 1.0, "global_radius_coef": 1.0,
             "glottal_lowpass_cutoff": 4000.0, "mouth_impedance_radius": 2.0},
    "female": {"vocal_tract_length": 15.0, "glottal_pulse_tn_min": 32.0, "glottal_pulse_tn_max": 32.0, "breathiness": 15.0,
               "global_nasal_radius_coef": 0.86, "global_radius_coef": 0.86, "min_glottal_loss": 2.0, "max_glottal_loss": 5.0,
               "mouth_impedance_radius": 1.71},
    "large_child": {"vocal_tract_length": 12.5, "breathiness": 25.0, "global_nasal_radius_coef": 0.71,
                    "global_radius_coef": 0.71, "mouth_impedance_radius": 1.43},
    "small_child": {"vocal_tract_length": 10.0, "breathiness": 30.0, "global_nasal_radius_coef": 0.57,
                    "global_radius_coef": 0.57, "mouth_impedance_radius": 1.14},
    "baby": {"vocal_tract_length": 7.5, "breathiness": 30.0, "global_nasal_radius_coef": 0.43, "global_radius_coef": 0.43,
             "mouth_impedance_radius": 0.86},
}

VOICE5_KEYS = sorted(set(VTM5_TXT) | set(_COMMON5) | set(VARIANTS5["male"]))


def default_voice5(variant="male"):
    """The shipped model-5 voice 5_male with one of its variants."""
    v = dict(VTM5_TXT)
    v.update(_COMMON5)
    v.update(VARIANTS5[variant])
    return v


def random_voice5(rng, base="male"):
    """Randomised model-5 voice: tract length (internal rate 60-90 kHz), glottal pulse shape, nasal radii, losses."""
    v = default_voice5(base)
    v["vocal_tract_length"] = float(rng.uniform(12.0, 17.5))
    v["glottal_pulse_tp"] = float(rng.uniform(30.0, 45.0))
    v["glottal_pulse_tn_min"] = float(rng.uniform(16.0, 28.0))
    v["glottal_pulse_tn_max"] = v["glottal_pulse_tn_min"] + float(rng.uniform(0.0, 12.0))
    for i in range(2, 8):
        v["nasal_radius_%d" % i] = _COMMON5["nasal_radius_%d" % i] * float(rng.uniform(0.7, 1.3))
    v["breathiness"] = float(rng.uniform(2.0, 30.0))
    v["max_glottal_loss"] = float(rng.uniform(1.0, 5.0))
    return v


def internal_rate5(voice):
    """Model 5's internal rate (VocalTractModel5.h:464-465): a double, 30 sections of one sample each."""
    length = voice["vocal_tract_length_offset"] + voice["vocal_tract_length"]
    length = min(max(length, 3.0), 30.0)
    c = 331.4 + (0.6 * voice["temperature"])
    return (c * 30 * 100.0) / length
