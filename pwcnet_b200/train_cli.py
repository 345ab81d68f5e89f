"""Training entry point: the reference's train.py (Trainer, train.py:24-172, and its argparse block 174-245) on the B200
compute path.

    python -m pwcnet_b200.train_cli -d SintelClean -dd /data/MPI-Sintel-complete -e 100 -b 4 --crop_shape 384 448

Same arguments, defaults and side effects as the reference: dataset + torch DataLoader (train.py:28-41; datasets from
pwcnet_b200/datahandler.py), PWCDCNet + multiscale loss + gamma * sum l2_loss(var) + Adam with the piecewise learning-rate
schedule (train.py:53-92), `--resume` from a checkpoint bundle, loss / EPE summaries every 1000 steps under
`logs/history_<time>/{train,val}` (JSON lines instead of TensorBoard events), a validation pass and
`./model/model_<epoch>.ckpt` (a tf.train.Saver-compatible bundle) per epoch, and the ExperimentSaver move at the end
(train.py:118-172).  Differences: no `input()` GPU prompt (use CUDA_VISIBLE_DEVICES / torchrun), no matplotlib figures
(`--visualize` is accepted and ignored: flow_utils.vis_flow_pyramid is out of scope), `--loss robust` is
rejected (the reference's multirobust_loss raises NameError, losses.py:45; only the multiscale loss has a backward path).
Under torchrun every rank trains on its own shard of each epoch and the flat gradient is all-reduced.
Images travel to the device as uint8; the `/255.0` of train.py:122 runs there."""
from __future__ import annotations

import argparse
import os
from datetime import datetime

import numpy as np
import torch


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument('-d', '--dataset', type=str, default='SintelClean', help='Target dataset, [SintelClean]')
    p.add_argument('-dd', '--dataset_dir', type=str, required=True, help='Directory containing target dataset')
    p.add_argument('-e', '--num_epochs', type=int, default=100, help='# of epochs [100]')
    p.add_argument('-b', '--batch_size', type=int, default=4, help='Batch size [4]')
    p.add_argument('-nw', '--num_workers', type=int, default=2, help='# of workers for data loading [2]')
    p.add_argument('--crop_type', type=str, default='random', help='Crop type for raw data [random]')
    p.add_argument('--crop_shape', nargs=2, type=int, default=[384, 448], help='Crop shape for raw data [384, 448]')
    p.add_argument('--resize_shape', nargs=2, type=int, default=None, help='Resize shape for raw data [None]')
    p.add_argument('--resize_scale', type=float, default=None, help='Resize scale for raw data [None]')
    p.add_argument('--num_levels', type=int, default=6, help='# of levels for feature extraction [6]')
    p.add_argument('--search_range', type=int, default=4, help='Search range for cost-volume calculation [4]')
    p.add_argument('--warp_type', default='bilinear', choices=['bilinear', 'nearest'], help='Warping protocol, [bilinear] or nearest')
    p.add_argument('--use-dc', dest='use_dc', action='store_true', help='Enable dense connection in optical flow estimator')
    p.add_argument('--no-dc', dest='use_dc', action='store_false', help='Disable dense connection in optical flow estimator')
    p.set_defaults(use_dc=False)
    p.add_argument('--output_level', type=int, default=4, help='Final output level for estimated flow [4]')
    p.add_argument('--loss', default='multiscale', choices=['multiscale', 'robust'], help='Loss function choice in [multiscale/robust]')
    p.add_argument('--lr', type=float, default=1e-4, help='Learning rate [1e-4]')
    p.add_argument('--lr_scheduling', dest='lr_scheduling', action='store_true', help='Enable learning rate scheduling')
    p.add_argument('--no-lr_scheduling', dest='lr_scheduling', action='store_false', help='Disable learning rate scheduling')
    p.set_defaults(lr_scheduling=True)
    p.add_argument('--weights', nargs='+', type=float, default=[0.32, 0.08, 0.02, 0.01, 0.005], help='Weights for each pyramid loss')
    p.add_argument('--gamma', type=float, default=0.0004, help='Coefficient for weight decay [4e-4]')
    p.add_argument('--epsilon', type=float, default=0.02, help='Small constant for robust loss [0.02]')
    p.add_argument('--q', type=float, default=0.4, help='Tolerance constant for outliear flow [0.4]')
    p.add_argument('-v', '--visualize', dest='visualize', action='store_true', help='accepted for compatibility; figures are out of scope')
    p.add_argument('--no-visualize', dest='visualize', action='store_false')
    p.set_defaults(visualize=True)
    p.add_argument('-r', '--resume', type=str, default=None, help='Learned parameter checkpoint file [None]')
    p.add_argument('--summary_every', type=int, default=1000, help='steps between training summaries (train.py:128) [1000]')
    p.add_argument('--max_steps', type=int, default=None, help='stop after this many optimisation steps (smoke tests)')
    return p


class Trainer(object):
    def __init__(self, args):
        self.args = args
        self._build_dataloader()
        self._build_graph()

    def _build_dataloader(self):
        from torch.utils import data
        from .datahandler import get_dataset
        a = self.args
        dset = get_dataset(a.dataset)
        data_args = {'dataset_dir': a.dataset_dir, 'origin_size': None, 'crop_type': a.crop_type, 'crop_shape': a.crop_shape,
                     'resize_shape': a.resize_shape, 'resize_scale': a.resize_scale}
        tset = dset(train_or_val='train', **data_args)
        vset = dset(train_or_val='val', **data_args)
        self.image_size = tset.image_size
        load_args = {'batch_size': a.batch_size, 'num_workers': a.num_workers, 'drop_last': True, 'pin_memory': True}
        self.num_batches = int(len(tset.samples) / a.batch_size)
        print(f'Found {len(tset.samples)} samples -> {self.num_batches} mini-batches')
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        sampler = None
        if self.world > 1:      # every rank draws its own shard of the epoch (data parallel)
            sampler = data.distributed.DistributedSampler(tset, num_replicas=self.world, rank=self.rank, shuffle=True, drop_last=True)
        self.tsampler = sampler
        self.tloader = data.DataLoader(tset, shuffle=sampler is None, sampler=sampler, **load_args)
        self.vloader = data.DataLoader(vset, shuffle=False, **load_args)

    def _build_graph(self):
        import torch.distributed as dist
        from . import PWCDCNet, Trainer as StepTrainer
        from .train import LR_BOUNDARIES
        from .utils import ExperimentSaver, SummaryWriter
        a = self.args
        if a.loss != 'multiscale':
            raise NotImplementedError("--loss robust: the reference's multirobust_loss raises NameError (losses.py:45) and no "
                                      "backward path exists for it here; use --loss multiscale")
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
        model = PWCDCNet(num_levels=a.num_levels, search_range=a.search_range, warp_type=a.warp_type, use_dc=a.use_dc,
                         output_level=a.output_level, name='pwcdcnet', device=f'cuda:{local}')
        # train.py:82-87: piecewise schedule, or a constant rate with --no-lr_scheduling
        bounds = LR_BOUNDARIES if a.lr_scheduling else []
        self.trainer = StepTrainer(model, lr=a.lr, gamma=a.gamma, weights=a.weights, lr_boundaries=bounds)
        if a.resume is not None:
            print(f'Loading learned model from checkpoint {a.resume}')
            self.trainer.load_state_dict(a.resume)
        logdir = 'logs/history_' + datetime.now().strftime('%Y-%m-%d-%H-%M')
        n = 1
        while os.path.exists(logdir):          # a second run within the same minute gets its own directory
            n += 1
            logdir = 'logs/history_' + datetime.now().strftime('%Y-%m-%d-%H-%M') + f'_{n}'
        self.is_chief = self.rank == 0
        if self.is_chief:
            self.twriter = SummaryWriter(logdir + '/train')
            self.vwriter = SummaryWriter(logdir + '/val')
            self.exp_saver = ExperimentSaver(logdir=logdir, parse_args=a)
            print(f'Graph building completed, histories are logged in {logdir}')
        self.logdir = logdir

    def train(self):
        from .pipeline import TrainStream
        from .utils import show_progress
        a, tr = self.args, self.trainer
        stream = TrainStream(tr, depth=2)
        done = False
        for e in range(a.num_epochs):
            if self.tsampler is not None:
                self.tsampler.set_epoch(e)
            it = iter(self.tloader)
            nxt = next(it, None)
            if nxt is not None:
                stream.submit(nxt[0][:, 0], nxt[0][:, 1], nxt[1])
            i = 0
            while nxt is not None:
                cur, nxt = nxt, next(it, None)
                if nxt is not None:                       # stage batch i+1 while step i runs
                    stream.submit(nxt[0][:, 0], nxt[0][:, 1], nxt[1])
                loss, loss_ms, epe = stream.step()
                g_step = tr.global_step
                i += 1
                if self.is_chief and g_step % a.summary_every == 0:                                     # train.py:128-132
                    l2, _, ep = tr.evaluate(cur[0][:, 0], cur[0][:, 1], cur[1])
                    self.twriter.add_summary({'loss/pwc': l2.item(), 'EPE/source': ep.item()}, g_step)
                    show_progress(e + 1, i, self.num_batches, loss=f'{l2.item():.4f}', epe=f'{ep.item():.4f}')
                if a.max_steps is not None and g_step >= a.max_steps:
                    done = True
                    if nxt is not None:
                        stream.discard()                   # the batch staged for the next step is not trained on
                    break
            # Validation (train.py:134-141): one summary per validation batch at the current step
            if self.is_chief:
                for images_val, flows_gt_val in self.vloader:
                    l2, _, ep = tr.evaluate(images_val[:, 0], images_val[:, 1], flows_gt_val)
                    self.vwriter.add_summary({'loss/pwc': l2.item(), 'EPE/source': ep.item()}, tr.global_step)
                os.makedirs('./model', exist_ok=True)
                tr.save(f'./model/model_{e + 1}.ckpt')                                                 # train.py:164-166
            if done:
                break
        if self.is_chief:
            self.twriter.close()
            self.vwriter.close()
            self.exp_saver.append(['./figure', './model'])
            self.exp_saver.save()
        return self.logdir


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.resize_scale is not None and not isinstance(args.resize_scale, (tuple, list)):
        args.resize_scale = (args.resize_scale, args.resize_scale)   # the reference unpacks a pair (flow.py:98)
    for key, item in vars(args).items():
        print(f'{key} : {item}')
    trainer = Trainer(args)
    return trainer.train()


if __name__ == '__main__':
    main()
