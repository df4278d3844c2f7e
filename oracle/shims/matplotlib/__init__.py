"""Import-time stand-in: Mgcod imports matplotlib in bin/visualizations.py but only plots with -m."""
