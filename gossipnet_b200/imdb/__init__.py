"""Image databases for the drivers (reference imdb/__init__.py:23-101).

`get_imdb(name, is_training)` returns the imdb dict the reference's train.py /
test.py consume, after the same preprocessing: optional single-class view
(`cfg.train.only_class`), images without detections dropped, and for training
the `max_num_detections` top-scoring cut and the appended mirror images.

Names: the reference's `coco_<year>_<split>` (needs data under
cfg.ROOT_DIR/data, none ships here) and `synthetic_<split>[_<images>x<dets>]`
[`_c<classes>`]: the SURVEY.md §8(d) generator as an imdb, which is what the
drivers run on in this repository.  CityPersons (`imdb/pal.py`) depends on a
`imdb.file_formats` module the reference itself does not ship: not provided.
"""
import os.path
import pickle
import re

import numpy as np

from gossipnet_b200 import synthetic
from gossipnet_b200.imdb import tools
from gossipnet_b200.imdb.coco import load_coco
from gossipnet_b200.nms_net.config import cfg

_imdbs = {}
for _year, _splits in (('2014', ('train', 'val', 'minival', 'valminusminival', 'minival2',
                                 'debug')),
                       ('2015', ('test', 'test-dev'))):
    for _split in _splits:
        _imdbs['coco_{}_{}'.format(_year, _split)] = (
            lambda split=_split, year=_year: load_coco(split, year))

_SYN = re.compile(r'^synthetic_(?P<split>[a-z]+)(_(?P<images>\d+)x(?P<dets>\d+))?(_c(?P<cls>\d+))?$')
_SPLIT_SEED = {'train': 0, 'val': 100000, 'minival': 200000, 'test': 300000}


def synthetic_imdb(name):
    """`synthetic_val_32x300_c80` -> 32 images x 300 detections, 80 classes."""
    m = _SYN.match(name)
    if m is None:
        raise KeyError(name)
    n_images = int(m.group('images') or 16)
    n_dets = int(m.group('dets') or 300)
    num_classes = int(m.group('cls') or 1)
    first = _SPLIT_SEED.get(m.group('split'), 400000)
    # class 1 is called 'person' so that experiments/coco_person/conf.yaml
    # (train.only_class: person) runs unchanged on a synthetic imdb
    classes = tuple(['__background__', 'person'] + ['class%d' % i for i in range(2, num_classes + 1)])
    roidb = []
    for i in range(n_images):
        img = synthetic.make_image(n_dets, num_classes, image_index=first + i)
        img.update(id=first + i, width=int(synthetic.CANVAS_W), height=int(synthetic.CANVAS_H),
                   filename='', flipped=False)
        roidb.append(img)
    return {'name': name, 'classes': classes,
            'class_to_ind': dict((c, i) for i, c in enumerate(classes)),
            'class_to_cat_id': dict((c, i) for i, c in enumerate(classes) if i > 0),
            'num_classes': num_classes, 'roidb': roidb}


def _load(name):
    if name in _imdbs:
        return _imdbs[name]()
    if name.startswith('synthetic_'):
        return synthetic_imdb(name)
    raise KeyError('unknown imdb {}'.format(name))


def get_imdb(name, is_training):
    """imdb/__init__.py:47-66: pickle cache under data/cache, then preprocessing.
    Synthetic imdbs are cheap to rebuild and are not cached."""
    result = None
    cache = os.path.join(cfg.ROOT_DIR, 'data', 'cache',
                         '{}_{}_imdb_cache.pkl'.format(name, cfg.train.detector))
    if os.path.exists(cache):
        print('reading {}'.format(cache))
        with open(cache, 'rb') as fp:
            result = pickle.load(fp)
    else:
        result = _load(name)
        if not name.startswith('synthetic_') and os.path.isdir(os.path.dirname(cache)):
            with open(cache, 'wb') as fp:
                pickle.dump(result, fp)
            print('wrote {}'.format(cache))
    (prepro_train if is_training else prepro_test)(result)
    return result


def _common_prepro(an_imdb):
    tools.print_stats(an_imdb)
    if cfg.train.only_class != '':
        print('dropping all classes but {}'.format(cfg.train.only_class))
        tools.only_keep_class(an_imdb, cfg.train.only_class)
        tools.print_stats(an_imdb)
    print('dropping images without detections')
    an_imdb['roidb'] = tools.drop_no_dets(an_imdb['roidb'])
    tools.print_stats(an_imdb)


def prepro_test(test_imdb):
    """imdb/__init__.py:69-79."""
    print('preparing test imdb')
    _common_prepro(test_imdb)
    print('done')


def prepro_train(train_imdb):
    """imdb/__init__.py:82-104."""
    print('preparing train imdb')
    _common_prepro(train_imdb)
    if cfg.train.max_num_detections > 0:
        print('dropping all but {} highest scoring detections'.format(
            cfg.train.max_num_detections))
        tools.drop_too_many_detections(train_imdb, cfg.train.max_num_detections)
        tools.print_stats(train_imdb)
    print('appending flipped images')
    train_imdb['roidb'] = tools.append_flipped(train_imdb['roidb'])
    train_imdb['avg_num_dets'] = tools.get_avg_batch_size(train_imdb)
    tools.print_stats(train_imdb)
    print('done')
