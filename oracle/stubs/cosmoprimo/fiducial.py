from .cosmology import Cosmology


def DESI(**kwargs):
    return Cosmology(**kwargs)
