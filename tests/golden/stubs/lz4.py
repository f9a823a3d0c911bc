def dumps(x):
    raise NotImplementedError("lz4 0.7.0 is not installable here")


loads = dumps
