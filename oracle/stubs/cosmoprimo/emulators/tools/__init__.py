__all__ = ['Emulator', 'PointEmulatorEngine', 'TaylorEmulatorEngine', 'MLPEmulatorEngine', 'Operation']


class Emulator(object):
    pass


class _BaseEngine(object):
    def __init__(self, *args, **kwargs):
        self.xoperations, self.yoperations = [], []


class PointEmulatorEngine(_BaseEngine):
    name = 'point'


class TaylorEmulatorEngine(_BaseEngine):
    name = 'taylor'


class MLPEmulatorEngine(_BaseEngine):
    name = 'mlp'


class Operation(object):
    pass
