"""Stand-ins for the counting cores of the reference's ``tasks/repeatability.py`` and ``tasks/MHA.py``."""
