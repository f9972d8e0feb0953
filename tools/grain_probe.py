"""Times the grain resampler (K6) on one 300 s track at +3 semitones: schedule on the host, kernel time
from the engine's CUDA-event profile (the host entry point also copies the result back)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import melonix_b200 as m  # noqa: E402
from melonix_b200 import hostlib as H  # noqa: E402
import signals as S  # noqa: E402

sec = float(sys.argv[1]) if len(sys.argv) > 1 else 300.0
x = S.vibrato_tone(sec, seed=9)
eng = m.Engine(0)
eng.upload_tracks([x])
gs, gl = H.grain_segment(x)
mk = [(10, 0.0, 0.0, 3.0), (x.size - 10, 0.0, 0.0, 3.0)]
s = H.export_schedule(x, 48000, mk, gs, gl)
for _ in range(2):
    eng.grain_render(0, s["gstart"], s["glen"], s["rate"], s["out_off"], s["next"], tail_zeros=s["tail_zeros"])
eng.profile_enable(True)
eng.profile_read()
reps = 5
for _ in range(reps):
    pcm, pcm16 = eng.grain_render(0, s["gstart"], s["glen"], s["rate"], s["out_off"], s["next"],
                                  tail_zeros=s["tail_zeros"])
ms, ln = eng.profile_read()["grain"]
ms /= ln
rate = float(np.mean(s["rate"]))
algo = pcm.size * (4 * rate + 4 + 2)
print(f"grain resampler: {pcm.size} output samples, {len(s['gstart'])} schedule rows, mean rate {rate:.4f}: "
      f"{ms * 1e3:.1f} us per launch, {algo / ms / 1e6:.0f} GB/s algorithmic (4*rate read + 4 + 2 written per sample)")
