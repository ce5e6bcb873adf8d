import functools, unittest
_tc = unittest.TestCase()
def make_decorator(func):
    def decorate(newfunc):
        newfunc = functools.wraps(func)(newfunc)
        return newfunc
    return decorate
def raises(*exceptions):
    def decorate(func):
        @functools.wraps(func)
        def newfunc(*a, **k):
            try:
                func(*a, **k)
            except exceptions:
                return
            raise AssertionError("%s() did not raise %s" % (func.__name__, exceptions))
        return newfunc
    return decorate
assert_raises = _tc.assertRaises
assert_equal = _tc.assertEqual
assert_equals = _tc.assertEqual
assert_not_equal = _tc.assertNotEqual
assert_true = _tc.assertTrue
assert_false = _tc.assertFalse
assert_almost_equal = _tc.assertAlmostEqual
assert_in = _tc.assertIn
assert_is_none = _tc.assertIsNone
assert_is_not_none = _tc.assertIsNotNone
assert_greater = _tc.assertGreater
assert_less = _tc.assertLess
assert_raises_regexp = _tc.assertRaisesRegex
def nottest(f):
    f.__test__ = False
    return f
def istest(f):
    f.__test__ = True
    return f
def with_setup(setup=None, teardown=None):
    def deco(f):
        return f
    return deco
def ok_(expr, msg=None):
    assert expr, msg
def eq_(a, b, msg=None):
    assert a == b, msg or "%r != %r" % (a, b)
