"""graphlearn_b200.ops - see the package README / DESIGN.md for the layer map."""
