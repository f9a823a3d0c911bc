"""Stand-in for the (absent) easydict package, used only by make_golden.py to import the reference."""


class EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = EasyDict(v) if isinstance(v, dict) else v

    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__
