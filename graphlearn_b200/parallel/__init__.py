"""graphlearn_b200.parallel - see the package README / DESIGN.md for the layer map."""
