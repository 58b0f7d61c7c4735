"""Stub: the reference imports matplotlib.pyplot at module top but the step path never plots."""
