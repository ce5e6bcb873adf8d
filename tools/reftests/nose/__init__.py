import unittest
SkipTest = unittest.SkipTest
from . import tools  # noqa
def run(*a, **k):
    return True
def runmodule(*a, **k):
    return True
