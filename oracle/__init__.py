"""oracle/ — TEST INFRASTRUCTURE ONLY (CPU restatement of the reference + build recipe of the real reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product (point-cloud-flow-matching_b200/) never does.
"""
