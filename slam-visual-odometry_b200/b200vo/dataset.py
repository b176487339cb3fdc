"""Image-sequence input -- the headless counterpart of the reference's main loop
(src/main.cpp:26-49 load_image_paths, :128-155 imread -> Frame -> process_frame).

  load_image_paths(dir)    png / jpg / jpeg (case-insensitive), lexicographic order, as the reference;
                           pgm / ppm are accepted too (what the C++ build can decode without OpenCV)
  timestamp_from_path      the file stem parsed as a number, else the frame index (main.cpp:146-151)
  read_image               BGR (H,W,3) u8 like cv::imread's default, or gray for 1-channel files;
                           decoding is done by Pillow (plumbing; the oracle / cv2 is never used here)
  run_directory            upload -> detect -> process_pairs in launch batches, chained on the host
"""
import os

import numpy as np

IMAGE_EXTENSIONS = (".png", ".jpg", ".jpeg", ".pgm", ".ppm")


def load_image_paths(directory):
    paths = []
    for name in os.listdir(directory):
        p = os.path.join(directory, name)
        if os.path.isfile(p) and os.path.splitext(name)[1].lower() in IMAGE_EXTENSIONS:
            paths.append(p)
    return sorted(paths)


def timestamp_from_path(path, index):
    stem = os.path.splitext(os.path.basename(path))[0]
    try:
        return float(stem)
    except ValueError:
        return float(index)


def read_image(path):
    from PIL import Image
    im = Image.open(path)
    if im.mode in ("L", "I;16", "1"):
        return np.ascontiguousarray(np.asarray(im.convert("L"), np.uint8))
    rgb = np.asarray(im.convert("RGB"), np.uint8)
    return np.ascontiguousarray(rgb[:, :, ::-1])          # BGR, the channel order cv::imread returns


def run_directory(directory, nfeatures=1200, batch_frames=64, device=0, **options):
    """Process every image of `directory` in order.  Returns (poses (n,12), records (n-1,), timestamps,
    paths); poses[i] = [R_wc | t_wc] of frame i, exactly what process_frame would have accumulated."""
    import b200vo
    paths = load_image_paths(directory)
    if not paths:
        raise FileNotFoundError("no images (%s) in %s" % (", ".join(IMAGE_EXTENSIONS), directory))
    first = read_image(paths[0])
    H, W = first.shape[:2]
    stamps = np.array([timestamp_from_path(p, i) for i, p in enumerate(paths)])
    B = max(2, min(batch_frames, len(paths)))
    recs = []
    with b200vo.Context(W, H, device=device, nfeatures=nfeatures, max_frames=B + 1, **options) as ctx:
        done, prev_slot = 0, 0
        while done < len(paths):
            n = min(B, len(paths) - done)
            imgs = [first if (done + k) == 0 else read_image(paths[done + k]) for k in range(n)]
            for im, p in zip(imgs, paths[done:done + n]):
                if im.shape[:2] != (H, W) or im.ndim != first.ndim:
                    raise ValueError("%s: size/channels differ from the first image" % p)
            slot0 = 0 if done == 0 else (prev_slot + 1) % (B + 1)
            ctx.upload_frames(np.stack(imgs), slot0)
            ctx.detect(slot0, n)
            p0, np_ = (slot0, n - 1) if done == 0 else (prev_slot, n)
            if np_ > 0:
                recs.append(ctx.process_pairs(p0, np_, done - (0 if done == 0 else 1)))
            prev_slot = (slot0 + n - 1) % (B + 1)
            done += n
    recs = np.concatenate(recs) if recs else np.zeros(0, b200vo.POSE_RECORD_DTYPE)
    poses = np.zeros((len(paths), 12))
    poses[0, :9] = np.eye(3).ravel()
    if len(recs):
        poses[1:] = b200vo.chain_trajectory(recs)
    return poses, recs, stamps, paths


def main(argv=None):
    import argparse
    import json
    from . import trajectory as T
    ap = argparse.ArgumentParser(description="headless VO over a directory of images (the reference's slam_vo_run without the GUI)")
    ap.add_argument("directory")
    ap.add_argument("--features", type=int, default=1200)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--out", default=None, help="write the trajectory in TUM format")
    a = ap.parse_args(argv)
    poses, recs, stamps, paths = run_directory(a.directory, a.features, a.batch)
    R, t = T.poses_to_Rt(poses)
    if a.out:
        T.save_tum(a.out, R, t, stamps)
    print(json.dumps({"frames": len(paths), "pairs_pose_ok": int(((recs["status"] & 4) != 0).sum()), "end_position": t[-1].round(4).tolist()}))


if __name__ == "__main__":
    main()
