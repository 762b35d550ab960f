"""CPU checker for polycracker_b200 -- TEST INFRASTRUCTURE ONLY (see py_oracle.py / oracle.c headers).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product package never does.
"""
