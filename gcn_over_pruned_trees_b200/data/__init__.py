"""Mirror of the reference's ``data`` package for the rows SURVEY.md 8f marks "next": the loader, device-resident."""
