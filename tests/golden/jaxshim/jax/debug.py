def callback(fn, *args, **kwargs):
    fn(*args, **kwargs)


def print(fmt, *args, **kwargs):  # noqa: A001
    import builtins

    builtins.print(fmt.format(*args, **kwargs))
