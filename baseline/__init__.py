"""Staging area of the unmodified reference (see stage_ref.sh).  Test / benchmark infrastructure only: nothing in the
product package imports from here."""
