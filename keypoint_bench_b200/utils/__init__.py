"""Same-named stand-ins for the reference's ``utils`` modules on the hot path
(utils/extracter.py, utils/matcher.py, utils/projection.py).  See INTEGRATION.md."""
