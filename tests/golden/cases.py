"""Golden model cases shared by make_golden.py (writer) and the tests (readers)."""

from tests import fixtures as fx

MODEL_CASES = {
    # name: (config, ref class, H, W, levels, B, rollout_step, seed)
    "tiny_33x64": ("tiny", "Aurora", 33, 64, fx.LEVELS13, 1, 0, 1),
    "tiny_lora_60x120_b2": ("tiny_lora", "Aurora", 60, 120, fx.LEVELS4, 2, 1, 2),
    "tiny_lora_all_step2": ("tiny_lora_all", "Aurora", 33, 64, fx.LEVELS4, 1, 2, 3),
    "tiny_lora_all_step3": ("tiny_lora_all", "Aurora", 33, 64, fx.LEVELS4, 1, 3, 3),
    "tiny_12h_stab_step0": ("tiny_12h_stab", "Aurora", 17, 32, fx.LEVELS4, 1, 0, 4),
    "tiny_12h_stab_step1": ("tiny_12h_stab", "Aurora", 17, 32, fx.LEVELS4, 1, 1, 4),
    "tiny_air_46x90": ("tiny_air", "AuroraAirPollution", 46, 90, fx.LEVELS13, 1, 0, 5),
    "tiny_air_46x90_step2": ("tiny_air", "AuroraAirPollution", 46, 90, fx.LEVELS13, 1, 2, 5),
    "tiny_wave_33x64": ("tiny_wave", "AuroraWave", 33, 64, fx.LEVELS4, 1, 0, 8),
    "tiny_wave_33x64_step1": ("tiny_wave", "AuroraWave", 33, 64, fx.LEVELS4, 1, 1, 8),
    "small_17x32": ("small", "AuroraSmallPretrained", 17, 32, fx.LEVELS4, 1, 0, 6),
}
