"""Sobol' analysis (drop-in for SALib.analyze.sobol)."""
