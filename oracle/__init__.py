"""oracle/ -- TEST INFRASTRUCTURE ONLY (see oracle/odtk_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product package must never do so."""
