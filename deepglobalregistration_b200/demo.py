"""Register two point-cloud files - the job of the reference's demo.py (:28-48), same flags:

    python -m deepglobalregistration_b200.demo --pcd0 a.ply --pcd1 b.ply --weights ckpt.pth

Reads PLY / KITTI .bin / 3DMatch .npz / .npy / text (io.py), prints the 4x4 pose that maps
pcd0 into pcd1's frame, optionally writes the moved cloud.  There is no download step (no
network) and no viewer."""
import argparse
import json

import numpy as np


def get_parser():
  ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
  ap.add_argument('--pcd0', required=True)
  ap.add_argument('--pcd1', required=True)
  ap.add_argument('--weights', required=True, help='reference checkpoint (state_dict, state_dict_inlier, config)')
  ap.add_argument('--clip_weight_thresh', type=float, default=0.05)     # config.py:63
  ap.add_argument('--no_icp', action='store_true', help='return the pre-ICP pose')
  ap.add_argument('--out', default=None, help='write pcd0 moved into the frame of pcd1 (PLY)')
  ap.add_argument('--json', action='store_true', help='print pose and diagnostics as one JSON line')
  return ap


def main(argv=None):
  args = get_parser().parse_args(argv)
  from . import io as dio
  from .core.deep_global_registration import DeepGlobalRegistration
  pcd0, pcd1 = dio.read_point_cloud(args.pcd0), dio.read_point_cloud(args.pcd1)
  config = argparse.Namespace(weights=args.weights, clip_weight_thresh=args.clip_weight_thresh,
                              verbose=not args.json)
  dgr = DeepGlobalRegistration(config)
  dgr.use_icp = not args.no_icp
  T01 = dgr.register(pcd0, pcd1)
  if args.json:
    print(json.dumps(dict(T=T01.tolist(), branch=dgr.last_branch, **dgr.last_info)))
  else:
    with np.printoptions(precision=6, suppress=True):
      print(T01)
  if args.out:
    dio.write_ply(args.out, pcd0.transform(T01).points, dtype='double')
  return T01


if __name__ == '__main__':
  main()
