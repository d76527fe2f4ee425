"""Drop-in for the reference's generate_gt.py: same flags, same `gt.log` files, with the overlap
ratios computed on the GPU (pointreggpt_b200.overlap).

    python generate_gt.py --dataset_name generated_dataset -start 0 -stop 10 --num_samples 2
"""
import argparse

from pointreggpt_b200.overlap import gather_gt, generate_gt

parser = argparse.ArgumentParser()
parser.add_argument('--dataset_name', default='generated_dataset', type=str, help='', required=True)
parser.add_argument('--start_scene_index', '-start', default=0, type=int, help='scenes index to start')
parser.add_argument('--stop_scene_index', '-stop', default=1, type=int, help='scenes index to stop')
parser.add_argument('--num_samples', default=2, type=int, help='sample numbers for each scene')
parser.add_argument('--disable_tqdm', action="store_true", help='disable tqdm')

if __name__ == "__main__":
    args = parser.parse_args()
    generate_gt(args.dataset_name, args.start_scene_index, args.stop_scene_index, args.num_samples)
    gather_gt(args.dataset_name, args.start_scene_index, args.stop_scene_index)
