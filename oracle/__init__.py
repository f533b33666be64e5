"""CPU oracle for the merPCR STS-search path -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; the product
package (merpcr_b200/) must never import this.
"""
