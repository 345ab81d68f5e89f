"""Bookkeeping helpers of the reference's training script (utils.py:9-53) -- progress line, config dump and the
`ExperimentSaver` that moves a run's artefacts into its log directory -- plus `SummaryWriter`, a JSON-lines stand-in for
the two `tf.summary.FileWriter`s of train.py:103-113 (TensorBoard event files need TensorFlow; the scalars and their
tags `loss/pwc`, `EPE/source` are the same)."""
from __future__ import annotations

import json
import sys
import time
from collections import OrderedDict
from datetime import datetime
from pathlib import Path


def show_progress(epoch, batch, batch_total, **kwargs):
    """utils.py:9-14."""
    fields = ''.join(f', {k}: {v}' for k, v in kwargs.items())
    sys.stdout.write(f'\r{epoch} epoch: [{batch}/{batch_total}{fields}]')
    sys.stdout.flush()


def save_config(config, filename=None):
    """utils.py:17-27."""
    if not isinstance(config, (dict, OrderedDict)):
        raise TypeError('arg config must be a dict or OrderedDict')
    if filename is None:
        filename = 'config_' + datetime.now().strftime('%Y-%m-%d-%H-%M') + '.json'
    with open(filename, 'w') as f:
        json.dump(OrderedDict(config), f, indent=4)
    print(f'Given config has been successfully saved to {filename}.')
    return filename


class ExperimentSaver:
    """utils.py:30-53: collects files / directories produced by a run and, on save(), renames them into `logdir`."""

    def __init__(self, logdir=None, parse_args=None):
        self.logdir = Path(logdir) if logdir is not None else Path('logs_' + datetime.now().strftime('%Y-%m-%d-%H-%M'))
        self.logdir.mkdir(parents=True, exist_ok=True)
        self.save_list = []
        if parse_args is not None:
            save_config(vars(parse_args), 'config.json')
            self.append('config.json')

    def append(self, file_or_dir_names):
        names = file_or_dir_names if isinstance(file_or_dir_names, list) else [file_or_dir_names]
        self.save_list.extend(Path(n) for n in names)

    def save(self):
        for path in self.save_list:
            if path.exists():                    # the reference raises when e.g. ./figure was never created
                path.rename(self.logdir / path.name)


class SummaryWriter:
    """`tf.summary.FileWriter(logdir)` + `add_summary(summary, step)` for scalars, as JSON lines in `<logdir>/scalars.jsonl`:
    {"step": ..., "wall_time": ..., "loss/pwc": ..., "EPE/source": ...}."""

    def __init__(self, logdir):
        self.logdir = Path(logdir)
        self.logdir.mkdir(parents=True, exist_ok=True)
        self._f = open(self.logdir / 'scalars.jsonl', 'a')

    def add_summary(self, scalars: dict, step: int):
        rec = {"step": int(step), "wall_time": time.time()}
        rec.update({k: float(v) for k, v in scalars.items()})
        self._f.write(json.dumps(rec) + '\n')
        self._f.flush()

    def close(self):
        self._f.close()
