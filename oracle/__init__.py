"""ORACLE: CPU restatements of the reference's read-path arithmetic (test infrastructure).

Importers allowed: tests/, __graft_entry__.smoke(), bench.py CPU legs.  The product
package (handwritten-ocr_b200/) never imports this package.
"""
