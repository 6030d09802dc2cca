import functools, warnings
def deprecate_func(*, since, package_name=None, removal_timeline=None, additional_msg=None, **kw):
    def deco(f):
        @functools.wraps(f)
        def w(*a, **k):
            warnings.warn(f"{f.__name__} deprecated since {since}. {additional_msg}", DeprecationWarning, stacklevel=2)
            return f(*a, **k)
        return w
    return deco
