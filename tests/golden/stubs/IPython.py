def embed(*a, **k):
    raise RuntimeError("IPython.embed() reached while generating goldens")
