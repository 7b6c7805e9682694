"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
   compute-sanitizer --tool racecheck python profiles/sanitize.py"""
import os
import sys


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcl_3dl_b200 import engine, synth  # noqa: E402

s = synth.scene(20_000, 70, 96, 40, seed=3)
for mapping in ("tuned", "group"):
    if mapping == "group":
        os.environ["MCL3DL_MAPPING"] = "group"
    for use_dda in (True, False):
        eng = engine.Engine((0,))
        lik = engine.LikParams(dist_weight=(1, 1, 5))
        beam = engine.beam_params_from_reference(num_points_default=40, use_raycast_using_dda=use_dda)
        eng.set_map(s["map"], lik, beam)
        r = eng.measure(s["particles"], s["lik"], s["beam"], s["origins"])
        st = eng.beam_status(s["particles"][:5], s["beam"], s["origins"])
        print(mapping, "dda" if use_dda else "kd", int(r["match_cnt"].sum()), int(r["n_hit"].sum()), st.shape)
        eng.close()
print("sanitize run complete")
