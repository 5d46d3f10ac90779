"""B200-native UAHN forward (CUAHN-VIO hot path).  See DESIGN.md."""
