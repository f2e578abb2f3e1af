"""Empty stand-in: the reference imports statsmodels in modules off the inference path."""
