"""ORACLE SUPPORT: attribute-dict stand-in for the `easydict` package the
reference's nms_net/config.py imports (not installed in this image)."""


class EasyDict(dict):
    def __init__(self, d=None, **kwargs):
        super(EasyDict, self).__init__()
        d = dict(d or {})
        d.update(kwargs)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, name, value):
        if isinstance(value, (list, tuple)):
            value = type(value)(self.__class__(x) if isinstance(x, dict) else x for x in value)
        elif isinstance(value, dict) and not isinstance(value, EasyDict):
            value = self.__class__(value)
        super(EasyDict, self).__setitem__(name, value)

    __setitem__ = __setattr__

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)
