"""ORACLE package — CPU restatements of the reference algorithms. Test infrastructure only."""
