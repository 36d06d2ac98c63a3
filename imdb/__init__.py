"""Drop-in import path: `import imdb`, `imdb.get_imdb(...)`, `imdb.tools...`
resolve to gossipnet_b200.imdb (the reference's drivers import it top-level)."""
import importlib
import sys

_impl = importlib.import_module('gossipnet_b200.imdb')
for _sub in ('tools', 'coco', 'detections'):
    sys.modules['imdb.' + _sub] = importlib.import_module('gossipnet_b200.imdb.' + _sub)
sys.modules[__name__] = _impl
