def __getattr__(name):
    raise RuntimeError("matplotlib is not available in this image (plots are outside the measured path)")
