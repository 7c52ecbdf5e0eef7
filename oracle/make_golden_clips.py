"""Small on-disk fixtures in the reference's clip / fit-result formats, cut from the reference's own
data files (5 AMASS frames, 4 Mixamo frames, one fit result), plus the values the reference's readers
produce for them (lib/model2video.py:527-531, lib/model2video_miaxmo.py:544-551, main.py:50-59).
Runs only in the build container (needs /root/reference).

Usage:  python oracle/make_golden_clips.py
"""
import os
import pickle

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/data"
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    d = np.load(os.path.join(REF, "amsass/09_05_poses.npz"))
    keep = [0, 1, 2, 70, 142]
    n0 = d["poses"].shape[0]
    np.savez_compressed(os.path.join(OUT, "amass_clip_fixture.npz"),
                        **{k: (d[k][keep] if d[k].ndim > 0 and d[k].shape[0] == n0 else d[k]) for k in d.files})
    # a whole real clip (143 frames, 156-D poses with hands, float32) for the 100k-frame replay of
    # BASELINE config 5: the tests tile it, exactly as SURVEY 8(d) prescribes
    np.savez_compressed(os.path.join(OUT, "amass_clip_09_05.npz"), poses=d["poses"].astype(np.float32),
                        trans=d["trans"].astype(np.float32), betas=d["betas"].astype(np.float32),
                        mocap_framerate=d["mocap_framerate"])
    m = pickle.load(open(os.path.join(REF, "mixamo/0007/result.pkl"), "rb"), encoding="iso-8859-1")
    n = 4
    mm = {"anim_len": n, "smpl_array": np.asarray(m["smpl_array"])[:n], "cam_array": np.asarray(m["cam_array"])[:n]}
    pickle.dump(mm, open(os.path.join(OUT, "mixamo_result_fixture.pkl"), "wb"), protocol=2)
    r = pickle.load(open(os.path.join(REF, "tests/test01/smplh.pkl"), "rb"), encoding="iso-8859-1")
    pickle.dump({k: np.asarray(v) for k, v in r.items()}, open(os.path.join(OUT, "fit_result_fixture.pkl"), "wb"), protocol=2)
    # what the reference's readers return for these files
    exp = dict(amass_poses=d["poses"][keep][:, :72], amass_trans=d["trans"][keep] - d["trans"][keep][0],
               mixamo_pose=np.asarray(m["smpl_array"])[:n].reshape(n, -1),
               fit_pose=np.asarray(r["spmlh_pose"]).reshape(-1, 3).astype("float64"),
               fit_shape=np.asarray(r["spmlh_shape"]).astype("float64"))
    np.savez_compressed(os.path.join(OUT, "clip_expected.npz"), **exp)


if __name__ == "__main__":
    main()
