import os, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import melonix_b200 as m
import signals as S
eng = m.Engine(0)
xs = [S.vibrato_tone(6.0, seed=21), S.vibrato_tone(4.3, seed=22)]
eng.upload_tracks(xs)
for semis in (-2.0, 3.0):
    r = m.semitone_ratio(semis)
    a = eng.pv_run(2048, 512, r, wave_mib=-1)
    a2 = eng.pv_run(2048, 512, r, wave_mib=-1)
    print(semis, "repeatable", [np.array_equal(u["y"], v["y"]) for u, v in zip(a, a2)])
    for chunk in ("", "24", "16", "64"):
        if chunk: os.environ["MLX_PV_CHUNK"] = chunk
        else: os.environ.pop("MLX_PV_CHUNK", None)
        for w in (-1, 1, 2, 5):
            b = eng.pv_run(2048, 512, r, wave_mib=w)
            res = []
            for u, v in zip(a, b):
                d = np.nonzero(u["y"] != v["y"])[0]
                res.append((d.size, int(d[0]) // 512 if d.size else -1, int(d[-1]) // 512 if d.size else -1, float(np.abs(u["y"] - v["y"]).max()), bool(np.array_equal(u["peak"], v["peak"]))))
            print(f"st={semis} chunk={chunk or 'def'} wave={w}: {res}")
