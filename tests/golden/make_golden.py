#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (needs /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports the reference's own modules (lib.model.rie, lib.model.Model, lib.camera.camera,
lib.train_val.trainer.eval_data_prepare), loads seeded synthetic weights
(ray3d_b200.synth, numpy PCG64 => reproducible on any box) with load_state_dict(strict=True)
and stores inputs plus fp32 / fp64 outputs as small .npz fixtures beside this script.
Weights are NOT stored (80-200 MB each); each fixture carries a float64 digest of the state
dicts so tests can prove the regenerated weights are the ones the reference saw.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("RAY3D_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from ray3d_b200.spec import NetSpec  # noqa: E402
from ray3d_b200 import synth  # noqa: E402

from lib.model.rie import RIEModel, RIETrajectoryModel  # noqa: E402  (reference)
from lib.camera.camera import CameraInfoPacket, normalize_screen_coordinates  # noqa: E402  (reference)

CASES = {
    # name: (spec kwargs, batch, resolution, uv kind)
    "h36m_s1_t27": (dict(num_joints=17, in_features=3, filter_widths=(3, 3, 3), stage=1), 3, 1000, "smooth"),
    "h36m_s3_t9": (dict(num_joints=17, in_features=3, filter_widths=(3, 3), stage=3), 4, 1000, "smooth"),
    "humaneva_s1_t9": (dict(num_joints=15, in_features=3, filter_widths=(3, 3), stage=1), 3, 1000, "uniform"),
    "h36mcross_s2_t9": (dict(num_joints=14, in_features=3, filter_widths=(3, 3), stage=2), 3, 1000, "smooth"),
    "rie_s1_t9_noembed": (dict(num_joints=17, in_features=2, filter_widths=(3, 3), stage=1, extrinsic_dim=0,
                               embed_dim=0), 3, 1000, "smooth"),
    "rie15_s3_t27": (dict(num_joints=15, in_features=2, filter_widths=(3, 3, 3), stage=3, extrinsic_dim=0,
                          embed_dim=0), 2, 1000, "smooth"),
    "h36m_s1_t81": (dict(num_joints=17, in_features=3, filter_widths=(3, 3, 3, 3), stage=1), 2, 1000, "smooth"),
    "h36m_s1_t243": (dict(num_joints=17, in_features=3, filter_widths=(3, 3, 3, 3, 3), stage=1), 2, 1000, "smooth"),
    "3dhp_s3_t243": (dict(num_joints=17, in_features=3, filter_widths=(3, 3, 3, 3, 3), stage=3), 2, 2048, "smooth"),
}


def build_reference(spec: NetSpec, sd_pos, sd_trj, dtype):
    kw = dict(filter_widths=list(spec.filter_widths), causal=False, dropout=0.2, latten_features=spec.latent,
              channels=spec.channels, dense=False, is_train=False, Optimize1f=True, stage=spec.stage,
              extrinsic_dim=spec.extrinsic_dim, embedd_dim=spec.embed_dim)
    pos = RIEModel(spec.num_joints, spec.in_features, spec.num_joints, **kw)
    trj = RIETrajectoryModel(spec.num_joints, spec.in_features, spec.num_joints, **kw)
    pos.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd_pos.items()}, strict=True)
    trj.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd_trj.items()}, strict=True)
    return pos.to(dtype).eval(), trj.to(dtype).eval()


def encode_with_reference(uv, cam, res):
    """Ray-encode every sequence with the reference's CameraInfoPacket (undistort=False)."""
    out = np.empty(uv.shape[:-1] + (3,), dtype=np.float64)
    for b in range(uv.shape[0]):
        fx, fy, cx, cy, pitch, height = (float(v) for v in cam[b])
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float64)
        c, s = np.cos(pitch), np.sin(pitch)
        # A camera pitched by `pitch` about x looking along world +y, z up: Rw2c rows = cam axes in world.
        R = np.array([[1, 0, 0], [0, -s, -c], [0, c, -s]], dtype=np.float64)
        t = -R @ np.array([[0.0], [0.0], [height]])
        pkt = CameraInfoPacket(P=None, K=K, R=R, t=t, dist_coeff=None, res_w=res, res_h=res, undistort=False)
        # replace the derived pitch by the table's pitch so both sides encode with the same angle
        pkt.cam_pitch_rad = pitch
        pkt.Rc2n, pkt.Tc2n = pkt.get_norm_coord_config()
        out[b] = pkt.get_cam_ray_given_uv(uv[b].astype(np.float64))
    return out


def main():
    torch.manual_seed(14)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    meta = {}
    for idx, (name, (kw, batch, res, kind)) in enumerate(CASES.items()):
        spec = NetSpec(**kw)
        sd_pos, sd_trj = synth.make_state_dicts(spec)
        uv, cam = synth.make_inputs(spec, batch, seed=1234 + idx, res=res, kind=kind)
        if spec.in_features == 3:
            x64 = encode_with_reference(uv, cam, res)
        else:
            x64 = normalize_screen_coordinates(uv.astype(np.float64), w=res, h=res)
        x = x64.astype(np.float32)                                     # trainer.py:298
        param = np.ascontiguousarray(cam[:, [5, 4]]) if spec.camera_embedding else np.zeros((batch, 2), np.float32)
        out = {}
        for tag, dtype in (("32", torch.float32), ("64", torch.float64)):
            pos, trj = build_reference(spec, sd_pos, sd_trj, dtype)
            with torch.no_grad():
                xt = torch.from_numpy(x).to(dtype)
                pt = torch.from_numpy(param).to(dtype)
                out["pos" + tag] = pos(xt, pt).numpy()
                out["trj" + tag] = trj(xt, pt).numpy()
            if tag == "32":
                keys_pos = [(k, list(v.shape)) for k, v in pos.state_dict().items()]
                keys_trj = [(k, list(v.shape)) for k, v in trj.state_dict().items()]
                named_pos = [k for k, _ in pos.named_parameters()]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), uv=uv, cam=cam, x64=x64, x=x, param=param,
                            digest=np.array([synth.state_digest(sd_pos), synth.state_digest(sd_trj)]), **out)
        meta[name] = dict(spec=dict(kw, filter_widths=list(kw["filter_widths"])), batch=batch, res=res, kind=kind,
                          seed=1234 + idx, keys_pos=keys_pos, keys_trj=keys_trj, named_params_pos=named_pos,
                          receptive_field=pos.receptive_field())
        err = np.linalg.norm(out["pos32"] - out["pos64"]) / np.linalg.norm(out["pos64"])
        print(f"{name:22s} pos {out['pos32'].shape} trj {out['trj32'].shape} fp32-vs-fp64 relerr {err:.2e}")

    # camera fixtures: real-looking extrinsics through the reference class (pitch/height derivation)
    rng = np.random.Generator(np.random.PCG64(77))
    cams = []
    for i in range(6):
        ang = rng.uniform(-1.0, 1.0, size=3)
        Rx = np.array([[1, 0, 0], [0, np.cos(ang[0]), -np.sin(ang[0])], [0, np.sin(ang[0]), np.cos(ang[0])]])
        Ry = np.array([[np.cos(ang[1]), 0, np.sin(ang[1])], [0, 1, 0], [-np.sin(ang[1]), 0, np.cos(ang[1])]])
        Rz = np.array([[np.cos(ang[2]), -np.sin(ang[2]), 0], [np.sin(ang[2]), np.cos(ang[2]), 0], [0, 0, 1]])
        R = Rx @ Ry @ Rz
        t = rng.uniform(-3.0, 3.0, size=(3, 1))
        fx, fy = rng.uniform(1000, 1500, size=2)
        cx, cy = rng.uniform(450, 600, size=2)
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]])
        pkt = CameraInfoPacket(P=None, K=K, R=R, t=t, dist_coeff=None, res_w=1000, res_h=1002, undistort=False)
        uv = rng.uniform(0, 1000, size=(5, 17, 2))
        cams.append(dict(K=K, R=R, t=t, uv=uv, ray=pkt.get_cam_ray_given_uv(uv), pitch=pkt.cam_pitch_rad,
                         height=float((-pkt.Rw2c.T @ pkt.Tw2c)[2][0]), Rc2n=pkt.Rc2n,
                         enc=pkt.encode_uv_with_intrinsic(uv),
                         norm=normalize_screen_coordinates(uv, w=1000, h=1002)))
    np.savez_compressed(os.path.join(HERE, "camera.npz"),
                        **{f"{k}{i}": np.asarray(c[k]) for i, c in enumerate(cams) for k in c})

    # lens undistortion through the reference class (cv2.undistortPoints, camera.py:412-441, 460-471)
    dist = np.array([-0.2075, 0.2478, -0.0014, -0.00098, -0.00307])          # H36M-like k1 k2 p1 p2 k3
    pk = CameraInfoPacket(P=None, K=cams[1]["K"], R=cams[1]["R"], t=cams[1]["t"], dist_coeff=dist, res_w=1000, res_h=1002, undistort=True)
    uvd = np.random.default_rng(77).uniform(0, 1000, size=(7, 17, 2))
    np.savez_compressed(os.path.join(HERE, "camera_undistort.npz"), K=cams[1]["K"], R=cams[1]["R"], t=cams[1]["t"], dist=dist, uv=uvd,
                        und=pk.undistort_point(uvd), pp_cam=pk.pp_cam, enc=pk.encode_uv_with_intrinsic(uvd),
                        ray=pk.get_cam_ray_given_uv(uvd), cv2_version=np.array(__import__("cv2").__version__))

    # evaluation tail: normalized2world (camera.py:401-410) + MPJPE family (lib/loss/loss.py) as evaluate_core uses them
    from lib.loss.loss import mpjpe, n_mpjpe, mean_velocity_error, p_mpjpe  # noqa: E402  (reference)
    pkt = CameraInfoPacket(P=None, K=cams[0]["K"], R=cams[0]["R"], t=cams[0]["t"], dist_coeff=None, res_w=1000, res_h=1002, undistort=False)
    pred = rng.standard_normal(size=(50, 1, 17, 3)).astype(np.float32)
    target = (pred + 0.05 * rng.standard_normal(size=pred.shape)).astype(np.float32)
    pw, tw = torch.from_numpy(pkt.normalized2world(pred)), torch.from_numpy(pkt.normalized2world(target))   # trainer.py:355-364
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), pred=pred, target=target, Rn2w=pkt.Rn2w, Tn2w=pkt.Tn2w,
                        pred_world=pw.numpy(), target_world=tw.numpy(), mpjpe=mpjpe(pw, tw).item(),
                        mrpe=mpjpe(pw[:, :, 0:1, :], tw[:, :, 0:1, :]).item(), n_mpjpe=n_mpjpe(pw, tw).item(),
                        mpjve=mean_velocity_error(pw.numpy().reshape(-1, 17, 3), tw.numpy().reshape(-1, 17, 3)),
                        p_mpjpe=p_mpjpe(pw.numpy().reshape(-1, 17, 3), tw.numpy().reshape(-1, 17, 3)),
                        K=cams[0]["K"], R=cams[0]["R"], t=cams[0]["t"])

    # sliding-window materialisation (trainer.py:47-58) -- imported lazily, it pulls in lib.loss etc.
    try:
        from lib.train_val.trainer import Trainer  # noqa: E402  (reference)
        seq = torch.from_numpy(rng.standard_normal(size=(1, 40, 17, 3)).astype(np.float32))
        win, _ = Trainer.eval_data_prepare(27, seq, None)
        np.savez_compressed(os.path.join(HERE, "windows.npz"), seq=seq.numpy(), win=win.numpy())
    except Exception as e:  # pragma: no cover
        print("eval_data_prepare fixture skipped:", e)

    # flip test-time augmentation: the reference's OWN Trainer.evaluate_core(flip_test=True) (trainer.py:283-405) on one
    # synthetic video -- eval_data_prepare, the mirrored second forward, un-mirroring, torch.mean, pos += trj,
    # normalized2world and the five metrics.  The trainer object is built without its constructor (it only wires
    # config/optimizer state evaluate_core never touches); the camera wrapper records what normalized2world receives.
    try:
        from lib.train_val.trainer import Trainer  # noqa: E402  (reference)
        trng = np.random.Generator(np.random.PCG64(4711))
        spec = NetSpec(num_joints=17, in_features=3, filter_widths=(3, 3, 3), stage=1)
        sd_pos, sd_trj = synth.make_state_dicts(spec)
        F, T = 12, spec.receptive_field
        uvv = (trng.uniform(0.3, 0.7, size=(F + T - 1, 17, 2)) * 1000)
        pkt = CameraInfoPacket(P=None, K=cams[2]["K"], R=cams[2]["R"], t=cams[2]["t"], dist_coeff=None, res_w=1000, res_h=1002, undistort=False)
        rays = pkt.get_cam_ray_given_uv(uvv)[None]                       # (1, F+T-1, 17, 3) float64, like the dataset holds it
        target = trng.standard_normal(size=(1, F, 17, 3))
        kps_left, kps_right = [4, 5, 6, 11, 12, 13], [1, 2, 3, 14, 15, 16]

        class RecordingCamera:
            def __init__(self, pk):
                self.pk, self.Rw2c, self.Tw2c, self.cam_pitch_rad, self.seen = pk, pk.Rw2c, pk.Tw2c, pk.cam_pitch_rad, []

            def normalized2world(self, pt):
                self.seen.append(np.array(pt, copy=True))
                return self.pk.normalized2world(pt)

        class OneVideo:
            def __init__(self, cam):
                self.cam = cam

            def next_epoch(self):
                yield self.cam, target.copy(), rays.copy()

        out = {}
        for tag, dtype in (("32", torch.float32), ("64", torch.float64)):
            pos, trj = build_reference(spec, sd_pos, sd_trj, dtype)
            if dtype == torch.float64:                                   # evaluate_core feeds float32 tensors: promote them at the module boundary
                pos.register_forward_pre_hook(lambda m, a: tuple(t.double() for t in a))
                trj.register_forward_pre_hook(lambda m, a: tuple(t.double() for t in a))
            tr = Trainer.__new__(Trainer)
            tr.pos_model_test, tr.trj_model_test = pos, trj
            tr.model_config, tr.data_config = {"TRAJECTORY_MODEL": True}, {"RAY_ENCODING": True}
            tr.kps_left, tr.kps_right, tr.receptive_field = kps_left, kps_right, T
            cam = RecordingCamera(pkt)
            e = tr.evaluate_core(OneVideo(cam), flip_test=True)
            out["pred" + tag] = cam.seen[0]                              # pos_tta + trj_tta, normalised frame (trainer.py:353-355)
            out["metrics" + tag] = np.array(e, dtype=np.float64)         # mm: mpjpe, p_mpjpe, n_mpjpe, mpjve, root
        np.savez_compressed(os.path.join(HERE, "evaluate_core_tta.npz"), uv=uvv, rays=rays[0].astype(np.float32), target=target[0].astype(np.float32),
                            param=np.array([(-pkt.Rw2c.T @ pkt.Tw2c)[2][0], pkt.cam_pitch_rad]).astype(np.float32),
                            kps_left=np.array(kps_left), kps_right=np.array(kps_right), Rn2w=pkt.Rn2w, Tn2w=pkt.Tn2w,
                            K=cams[2]["K"], R=cams[2]["R"], t=cams[2]["t"], **out)
    except Exception as e:  # pragma: no cover
        print("evaluate_core flip-TTA fixture skipped:", repr(e))
        raise

    with open(os.path.join(HERE, "meta.json"), "w") as f:
        json.dump(meta, f, indent=0)
    print("wrote", len(CASES), "cases to", HERE)


if __name__ == "__main__":
    main()
