"""TEST INFRASTRUCTURE ONLY.

CPU oracle of the SVG-IR splatting + shading hot path (see the headers of svgss_oracle.c and
shading_oracle.py for the reference file:line each function restates). Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product (svg-ir_b200/) never does.
"""
