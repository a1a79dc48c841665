"""graphlearn_b200.data - see the package README / DESIGN.md for the layer map."""
