"""Empty-shell stand-in for the third-party `lsstypes` package (absent here).  Test infrastructure only."""


class CovarianceMatrix(object):
    pass


class ObservableTree(object):
    pass


class WindowMatrix(object):
    pass


class ObservableLeaf(object):
    pass


def read(*args, **kwargs):
    raise NotImplementedError('lsstypes stub')
