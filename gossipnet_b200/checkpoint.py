"""Checkpoint naming, resume and best-model bookkeeping of the reference
(train.py:40-61, 274-304, 337-345) for the flat-parameter trainer.

The reference saves with `tf.train.Saver(max_to_keep=None).save(sess, 'gnet',
global_step=it)` into the working directory: files `gnet-<it>.*` plus a
`checkpoint` state file whose `model_checkpoint_path` names the newest one, which
`--resume` reads back (`tf.train.get_checkpoint_state('./')`), continuing at
`global_step + 1`.  Here a checkpoint is ONE file `gnet-<it>` (torch.save of
the named parameter tensors, the optimizer slots and the step), the state file
has the same two-key text format, and `ModelManager` keeps the (iteration, mAP,
file) list and the `gnet_best` symlink.
Parameter names are the reference's TF variable names ([in, out] weights), so a
checkpoint is also a plain name -> array dictionary.
"""
import os

import torch

STATE_FILE = 'checkpoint'


def checkpoint_name(prefix, global_step):
    return '{}-{}'.format(prefix, int(global_step))


def _write_state(directory, newest, all_paths):
    with open(os.path.join(directory, STATE_FILE), 'w') as fp:
        fp.write('model_checkpoint_path: "{}"\n'.format(newest))
        for p in all_paths:
            fp.write('all_model_checkpoint_paths: "{}"\n'.format(p))


def get_checkpoint_state(directory='./'):
    """-> {'model_checkpoint_path': str, 'all_model_checkpoint_paths': [str]} or None."""
    path = os.path.join(directory, STATE_FILE)
    if not os.path.exists(path):
        return None
    state = {'model_checkpoint_path': None, 'all_model_checkpoint_paths': []}
    with open(path) as fp:
        for line in fp:
            key, _, val = line.partition(':')
            val = val.strip().strip('"')
            if key.strip() == 'model_checkpoint_path':
                state['model_checkpoint_path'] = val
            elif key.strip() == 'all_model_checkpoint_paths':
                state['all_model_checkpoint_paths'].append(val)
    return state if state['model_checkpoint_path'] else None


class Saver(object):
    """max_to_keep=None: every checkpoint is kept (train.py:288)."""

    def __init__(self, directory='./'):
        self.directory = directory
        prev = get_checkpoint_state(directory)
        self._paths = list(prev['all_model_checkpoint_paths']) if prev else []

    def save(self, trainer, prefix, global_step):
        """-> the path written (`<dir>/<prefix>-<global_step>`)."""
        name = checkpoint_name(prefix, global_step)
        path = os.path.join(self.directory, name)
        net = trainer.net
        blob = {
            'format': 'gossipnet_b200-checkpoint-1',
            'global_step': int(global_step),
            'optimizer': trainer.optimizer,
            'optimizer_steps': int(trainer.global_step),
            'variables': dict((k, v.detach().cpu().clone()) for k, v in net.state_dict().items()),
            'slot1': trainer.state1.detach().cpu(),
            'slot2': trainer.state2.detach().cpu(),
        }
        tmp = path + '.tmp'
        torch.save(blob, tmp)
        os.replace(tmp, path)          # never leave a half-written newest checkpoint
        if name not in self._paths:
            self._paths.append(name)
        _write_state(self.directory, name, self._paths)
        return path

    def restore(self, trainer, path):
        """Loads parameters + optimizer slots; returns the iteration to continue
        at (train.py:303-304: global_step + 1)."""
        blob = torch.load(path, map_location='cpu', weights_only=True)   # tensors + plain values only
        if blob.get('format') != 'gossipnet_b200-checkpoint-1':
            raise ValueError('{} is not a gossipnet_b200 checkpoint'.format(path))
        trainer.net.load_state_dict(blob['variables'])
        dev = trainer.state1.device
        trainer.state1.copy_(blob['slot1'].to(dev))
        trainer.state2.copy_(blob['slot2'].to(dev))
        trainer.global_step = int(blob['optimizer_steps'])
        return int(blob['global_step']) + 1

    def restore_latest(self, trainer):
        state = get_checkpoint_state(self.directory)
        if state is None:
            raise IOError('no checkpoint state file in {}'.format(self.directory))
        return self.restore(trainer, os.path.join(self.directory, state['model_checkpoint_path']))


def load_variables(path):
    """name -> tensor dictionary of a checkpoint (what test.py's restorer needs)."""
    blob = torch.load(path, map_location='cpu', weights_only=True)   # tensors + plain values only
    return blob['variables']


class ModelManager(object):
    """train.py:40-61."""

    def __init__(self):
        self.models = []

    def add(self, global_iter, ap, model_file):
        self.models.append((global_iter, ap, model_file))

    def best(self):
        return max((ap, model_file) for _, ap, model_file in self.models)

    def print_summary(self):
        _, best_file = self.best()
        print('{:10s}  {:6s}'.format('Iteration', 'mAP'))
        for it, ap, model_file in self.models:
            print('{:10d}  {:6.1f}{}'.format(it, ap, '  (best)' if model_file == best_file else ''))

    def write_link_to_best(self, link):
        _, best_file = self.best()
        print('writing symlink {} -> {}'.format(link, best_file))
        if os.path.lexists(link):
            os.remove(link)
        os.symlink(best_file, link)
