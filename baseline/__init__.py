"""Reference arm support: locating the vendored, unmodified songlei00/bbo (``baseline/_ref``, git-ignored,
created by ``baseline/vendor_reference.py``) and making it importable.  Test / benchmark infrastructure only:
nothing under ``bbo_b200/`` imports this package."""
