"""Sobol'/Saltelli sample generation (drop-ins for SALib.sample.sobol / SALib.sample.saltelli)."""
