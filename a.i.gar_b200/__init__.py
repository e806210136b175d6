"""B200-native batched agar.io env step (host side).  See DESIGN.md."""
