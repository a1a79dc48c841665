"""graphlearn_b200.engine - see the package README / DESIGN.md for the layer map."""
