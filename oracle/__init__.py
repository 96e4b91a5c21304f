"""CPU oracle for the map->GvdGraph hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing under active-orchard-slam_b200/ does.
"""
