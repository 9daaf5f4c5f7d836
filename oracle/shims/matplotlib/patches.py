def __getattr__(name):
    raise NotImplementedError('matplotlib stand-in: %s is not available' % name)
