"""Command-line front end (mirror of SALib.scripts.salib for the Sobol' path)."""
