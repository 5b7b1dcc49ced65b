"""Empty stand-in: the reference imports matplotlib.pyplot (sim_plain.py:3) but
its only plotting routine returns immediately (sim_plain.py:233-234)."""
