def flatarray(*args, **kwargs):
    def deco(func):
        return func
    if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]
    return deco


def addproperty(*attrs):
    def deco(cls):
        return cls
    return deco
