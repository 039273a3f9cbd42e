"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.  CPU oracle for the Jalla hot path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; jalla_b200/ never does."""
