def from_pypower(*args, **kwargs):
    raise NotImplementedError('lsstypes stub')


def from_pycorr(*args, **kwargs):
    raise NotImplementedError('lsstypes stub')
