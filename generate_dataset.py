"""Drop-in for the reference's generate_dataset.py (GD:1-63): same flags, same hard-coded
model / sampler hyper-parameters, same output layout -- running on the B200-native path.

    python generate_dataset.py --resume official [-start 0 -stop 10 --num_samples 1]
    torchrun --nproc-per-node 8 generate_dataset.py --resume official -start 0 -stop 10000

Extra (optional) flags: --data_root (3DMatch RGB-D train tree; it must exist) or --synthetic (seeded
synthetic source frames -- no dataset can be downloaded offline), --random_init (skip checkpoint
loading), --seed (base seed: scene k's pose and noise depend on (seed, k) only, so the dataset is the
same for any world size / batch size / -start -stop split), --sampling_timesteps / --batch_size.
"""
import argparse
import os

import torch

from pointreggpt_b200.diffusion import GaussianDiffusion
from pointreggpt_b200.generator import Generator
from pointreggpt_b200.nets import MaskUnet, Unet
from pointreggpt_b200 import dist as pdist

parser = argparse.ArgumentParser()
parser.add_argument('--resume', default=None, type=str, help='checkpoint to load', required=True)
parser.add_argument('--dataset_name', default='generated_dataset', type=str, help='')
parser.add_argument('--start_scene_index', '-start', default=0, type=int, help='scenes index to start')
parser.add_argument('--stop_scene_index', '-stop', default=1, type=int, help='scenes index to stop')
parser.add_argument('--num_samples', default=1, type=int, help='sample numbers for each scene')
parser.add_argument('--data_root', default='/path/to/3DMatch-RGBD/train', type=str)
parser.add_argument('--synthetic', action='store_true', help='seeded synthetic source frames instead of 3DMatch')
parser.add_argument('--random_init', action='store_true', help='seeded random weights, no checkpoint')
parser.add_argument('--seed', default=0, type=int, help='base seed of the per-scene pose / noise streams')
parser.add_argument('--sampling_timesteps', default=250, type=int)
parser.add_argument('--batch_size', default=4, type=int)
parser.add_argument('--device_batch', default=None, type=int,
                    help='scenes per pass through the GPU (default 32; results do not depend on it)')
args = parser.parse_args()

if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not torch.distributed.is_initialized():
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    torch.distributed.init_process_group("nccl")

if args.random_init:
    torch.manual_seed(0)          # weight construction only; sampling noise is keyed per scene (--seed)
model = Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1)
diffusion = GaussianDiffusion(model, image_size=256, timesteps=1000,
                              sampling_timesteps=args.sampling_timesteps, loss_type='l1',
                              objective='pred_x0', beta_schedule='sigmoid', ddim_sampling_eta=1.0,
                              is_ddnm_sampling=True)
if args.synthetic:
    folder = "synthetic"
elif os.path.isdir(args.data_root):
    folder = args.data_root
else:
    raise FileNotFoundError("--data_root %r does not exist (use --synthetic for the synthetic source "
                            "frames)" % args.data_root)
generator = Generator(diffusion, folder, batch_size=args.batch_size, device_batch=args.device_batch, ema_decay=0.995,
                      results_folder='./successive_ddnm_diffusion_results',
                      samples_folder='./{}/data'.format(args.dataset_name), amp=False)
depth_correction = MaskUnet(dim=64, dim_mults=(1, 2, 4, 8))
if args.random_init:
    with torch.no_grad():
        depth_correction.final_conv[0].bias.fill_(8.0)
else:
    generator.load("{}".format(args.resume))
if torch.distributed.is_initialized():
    pdist.broadcast_weights([generator.ema.ema_model, depth_correction.to(generator.device)], src=0)
n = generator.generate(start_scene_index=args.start_scene_index, stop_scene_index=args.stop_scene_index,
                       num_samples=args.num_samples, has_refine_step=False,
                       depth_correction=depth_correction, base_seed=args.seed)
if generator.rank == 0:
    print("generated %d scenes into ./%s/data" % (n, args.dataset_name))
