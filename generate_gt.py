"""Command-line front end with the flags of the reference's generate_gt.py; writes the same per-scene
`gt.log` files and the gathered `metadata/gt.log`.  The overlap ratios come from the GPU radius
search in pointreggpt_b200.overlap.

    python generate_gt.py --dataset_name generated_dataset -start 0 -stop 10 --num_samples 2
"""
import argparse

from pointreggpt_b200 import overlap


def cli():
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("--dataset_name", required=True, help="folder that holds data/scene-XXXXXX")
    for long_name, short, default, text in (
            ("--start_scene_index", "-start", 0, "first scene (inclusive)"),
            ("--stop_scene_index", "-stop", 1, "last scene (exclusive)")):
        ap.add_argument(long_name, short, type=int, default=default, help=text)
    ap.add_argument("--num_samples", type=int, default=2, help="clouds per scene; all pairs are rated")
    ap.add_argument("--disable_tqdm", action="store_true", help="accepted for compatibility; no progress bar here")
    return ap.parse_args()


if __name__ == "__main__":
    a = cli()
    overlap.generate_gt(a.dataset_name, a.start_scene_index, a.stop_scene_index, a.num_samples)
    overlap.gather_gt(a.dataset_name, a.start_scene_index, a.stop_scene_index)
