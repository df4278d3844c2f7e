def install_mp_handler(logger=None):
    """no-op stand-in (bin/multiprocess_mgcod.py:15,134)"""
