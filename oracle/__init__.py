"""TEST INFRASTRUCTURE ONLY: CPU oracle of the Scan2Cap hot path (see oracle/README.md).

Nothing under scan2cap_b200/ may import this package.
"""
