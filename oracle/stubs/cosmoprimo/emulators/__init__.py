from . import tools


class Samples(object):
    pass
