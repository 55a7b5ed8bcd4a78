"""Empty stand-in for matplotlib (TEST INFRASTRUCTURE ONLY): lets the reference's GraceRIGV3.py be
imported in the build container for golden-vector generation; nothing is ever plotted."""
