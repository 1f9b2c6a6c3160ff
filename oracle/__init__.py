"""ORACLE package: CPU restatements of the reference's scoring path.  Test infrastructure only —
imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs, never by the product
package `law_of_vision_representation_in_mllms_b200`."""
