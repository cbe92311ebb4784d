"""Entry point with the reference's flags and dense-method selection
(reference: src/run_reconstruction.py:24-186).

    python -m src.run_reconstruction <dataset> [--mvs | --stereo] [--fast] [--output DIR] [--poses FILE]

Only Stage 4 (dense) is implemented here.  Stages 1-3 (sparse SfM: OpenCV features, matching,
PnP -- reference src/core/sfm_pipeline.py) are out of scope of this engine, so camera poses come
from one of:
  * `--poses poses.npz`  (keys: `R` (n,3,3), `t` (n,3), `files` (n,) image names relative to the
    dataset directory, optional `points` (m,3) sparse cloud) -- known poses, e.g. from a previous
    sparse run or a synthetic scene;
  * the reference's own `SfMPipeline`, if its `src` directory is given in $REF_SRC (it is imported
    and run unmodified; nothing of it is vendored here).
Multi-GPU: `torchrun --nproc-per-node N -m src.run_reconstruction ...` shards the reference views over N GPUs (rank 0
writes the files).  Extra flags `--arith`, `--schedule`, `--consistency` select this engine's opt-in modes (defaults =
the reference's behaviour).
Precedence mvs > stereo, skipped with --fast or fewer than 3 poses; outputs `dense_mvs.ply` /
`dense_stereo.ply` -- all as in the reference (:118-172).  `--dense` (CPU SIFT matching,
reference src/core/dense.py) is not part of the GPU path and is rejected.
"""
import argparse
import os
import sys
from pathlib import Path

src_dir = Path(__file__).resolve().parent
pkg_dir = src_dir.parent
for p in (str(src_dir), str(pkg_dir)):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

from core.camera import CameraPose, load_calibration  # noqa: E402
from core.utils import save_cameras_ply, save_ply  # noqa: E402


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description='3D Reconstruction from Images (B200 dense stage)')
    p.add_argument('dataset', help='Dataset name (subfolder in data/samples/)')
    p.add_argument('--max-images', type=int, default=None, help='Maximum number of images to process')
    p.add_argument('--dense', action='store_true', help='Dense SIFT matching (CPU; not available in this engine)')
    p.add_argument('--stereo', action='store_true', help='Plane sweep stereo (fast, basic quality)')
    p.add_argument('--mvs', action='store_true', help='PatchMatch MVS (best quality, GPU accelerated)')
    p.add_argument('--combined', action='store_true', help='Combined stereo + dense (deprecated)')
    p.add_argument('--fast', action='store_true', help='Fast mode: sparse only, reduced resolution')
    p.add_argument('--neural', action='store_true', help='Use LightGlue neural matcher (sparse stage option)')
    p.add_argument('--output', type=str, default=None, help='Output directory')
    p.add_argument('--poses', type=str, default=None, help='Known poses (.npz) instead of running sparse SfM')
    p.add_argument('--calibration', type=str, default=None, help='calibration_data.npz (keys mtx, dist)')
    p.add_argument('--scale', type=float, default=0.25, help='Working resolution scale (reference: 0.25)')
    # not in the reference (defaults = the reference's behaviour): opt-in modes of this engine
    p.add_argument('--arith', choices=['parity', 'fast'], default='parity',
                   help="'parity' = bit-exact with the reference (default); 'fast' = re-associated arithmetic, NOT bit-exact")
    p.add_argument('--schedule', choices=['reference', 'redblack', 'viewprop'], default='reference',
                   help='PatchMatch propagation schedule; redblack / viewprop are extensions without a reference counterpart')
    p.add_argument('--consistency', choices=['photometric', 'geometric'], default='photometric',
                   help="'geometric' adds the cross-view depth-map check (extension)")
    return p


def init_distributed():
    """Under `torchrun --nproc-per-node N` the dense stage shards the reference views over N GPUs (one process per GPU,
    NCCL); rank 0 returns / writes the cloud.  Without torchrun this is a no-op (single GPU)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
        return 0, 1
    import torch
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    return dist.get_rank(), world


def load_known_poses(path: Path, image_dir: Path, max_images=None):
    import cv2 as cv
    z = np.load(str(path), allow_pickle=False)
    files = [str(f) for f in z['files']][:max_images]
    images, poses = [], {}
    for i, name in enumerate(files):
        img = cv.imread(str(image_dir / name))
        if img is None:
            raise FileNotFoundError(image_dir / name)
        images.append({'path': str(image_dir / name), 'image': img, 'gray': cv.cvtColor(img, cv.COLOR_BGR2GRAY)})
        poses[i] = CameraPose(R=z['R'][i].astype(np.float64), t=z['t'][i].astype(np.float64).ravel())
    points = z['points'] if 'points' in z.files else np.zeros((0, 3))
    colors = np.full((len(points), 3), 200, np.uint8)
    return images, poses, points, colors


def run_reference_sfm(calibration_path, image_dir, args):
    ref_src = os.environ.get('REF_SRC')
    if not ref_src or not (Path(ref_src) / 'core' / 'sfm_pipeline.py').exists():
        print("ERROR: sparse SfM is outside this engine. Pass --poses poses.npz, or set REF_SRC to the "
              "reference's src directory to run its SfMPipeline.")
        sys.exit(1)
    import importlib.util
    spec = importlib.util.spec_from_file_location('refcore', str(Path(ref_src) / 'core' / '__init__.py'),
                                                  submodule_search_locations=[str(Path(ref_src) / 'core')])
    refcore = importlib.util.module_from_spec(spec)
    sys.modules['refcore'] = refcore
    spec.loader.exec_module(refcore)
    pipeline = refcore.SfMPipeline(str(calibration_path), fast_mode=args.fast, neural_mode=args.neural)
    points, colors, poses = pipeline.reconstruct(str(image_dir), args.max_images)
    pose_dict = {}
    for idx, pose in poses.items():
        R, t = (pose.R, pose.t) if hasattr(pose, 'R') else pose
        pose_dict[idx] = CameraPose(R=np.asarray(R), t=np.asarray(t).ravel())
    return pipeline.images, pose_dict, points, colors


def main(argv=None):
    args = build_parser().parse_args(argv)
    rank, world = init_distributed()
    device = f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}" if world > 1 else None
    project_root = src_dir.parent
    calibration_path = Path(args.calibration) if args.calibration else (
        project_root / 'src' / 'calibration' / 'calibration_results' / 'calibration_data.npz')
    image_dir = project_root / 'data' / 'samples' / args.dataset
    if not image_dir.exists() and Path(args.dataset).exists():
        image_dir = Path(args.dataset)
    output_dir = Path(args.output) if args.output else image_dir / 'reconstruction'
    output_dir.mkdir(parents=True, exist_ok=True)
    if not calibration_path.exists():
        print(f"ERROR: Calibration file not found: {calibration_path}")
        print("Run camera calibration first!")
        sys.exit(1)
    if not image_dir.exists():
        print(f"ERROR: Image directory not found: {image_dir}")
        sys.exit(1)
    if args.dense and not (args.mvs or args.stereo):
        print("ERROR: --dense (CPU SIFT matching) is not part of the GPU dense stage; use --mvs or --stereo.")
        sys.exit(1)

    print("=" * 60)
    print("3D RECONSTRUCTION")
    print("=" * 60)
    print(f"Dataset: {args.dataset}")
    print(f"Images: {image_dir}")
    print(f"Output: {output_dir}")
    mode = (["FAST"] if args.fast else []) + ["LightGlue" if args.neural else "SIFT"]
    mode.append("PatchMatch MVS" if args.mvs else "Plane Sweep" if args.stereo else "Sparse only")
    print(f"Mode: {' + '.join(mode)}")
    print()

    if args.poses:
        images, pose_dict, points, colors = load_known_poses(Path(args.poses), image_dir, args.max_images)
    else:
        images, pose_dict, points, colors = run_reference_sfm(calibration_path, image_dir, args)
    if rank == 0:
        save_ply(points, colors, str(output_dir / 'sparse.ply'))
        save_cameras_ply(pose_dict, str(output_dir / 'cameras.ply'))

    run_dense = (args.mvs or args.stereo) and not args.fast
    if run_dense and len(pose_dict) >= 3:
        camera = load_calibration(str(calibration_path))
        if args.mvs:
            print("\n" + "=" * 60)
            print("Starting PatchMatch MVS...")
            print("=" * 60)
            from core.mvs_patchmatch import PatchMatchMVS
            mvs = PatchMatchMVS(camera, scale=args.scale, num_iterations=3, min_views=3, arith=args.arith,
                                schedule=args.schedule, consistency=args.consistency, device=device)
            dense_points, dense_colors = mvs.reconstruct(images, pose_dict,
                                                         sparse_points=points if len(points) else None)
            if rank == 0 and len(dense_points) > 0:
                dense_output = output_dir / 'dense_mvs.ply'
                save_ply(dense_points, dense_colors, str(dense_output))
                print(f"Saved {len(dense_points):,} points to {dense_output}")
        else:
            from core.dense_stereo import DenseStereoReconstructor
            stereo = DenseStereoReconstructor(camera, scale=args.scale, arith=args.arith, device=device)
            dense_points, dense_colors = stereo.reconstruct(images, pose_dict, max_pairs=30)
            if rank == 0 and len(dense_points) > 0:
                dense_output = output_dir / 'dense_stereo.ply'
                save_ply(dense_points, dense_colors, str(dense_output))
                print(f"Saved {len(dense_points):,} points to {dense_output}")

    print("\n" + "=" * 60)
    print("DONE!")
    print("=" * 60)
    print(f"Results saved to: {output_dir}")
    print(f"  - sparse.ply: {len(points):,} points")
    if args.mvs:
        print("  - dense_mvs.ply: PatchMatch MVS cloud")
    elif args.stereo:
        print("  - dense_stereo.ply: Plane sweep cloud")
    print(f"  - cameras.ply: {len(pose_dict)} camera positions")
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
